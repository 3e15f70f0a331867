"""Golden fixtures (tests/golden/).
CPU: the oracle replays the reference's known-answer vectors (reference_kats.json, each entry
cites the reference test it was transcribed from) and still reproduces the committed .npz files.
GPU: the CUDA path, through the C ABI, reproduces the committed .npz files WITHOUT the oracle."""
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
O = make_golden.O

with open(os.path.join(GOLD, "reference_kats.json")) as fh:
    KATS = json.load(fh)["kats"]


@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_reference_kat(kat):
    fn = kat["fn"]
    if fn == "top_k_indices":
        assert O.top_k_indices(kat["scores"], kat["k"]) == kat["expect"]
    elif fn == "top_k_indices_heap":
        assert O.top_k_indices(kat["scores"], kat["k"], heap=True) == kat["expect"]
    elif fn == "merge_search_results":
        ids, dist = O.merge_search_results(kat["ids"], kat["dist"], kat["k"])
        assert ids == kat["expect_ids"]
        assert dist == [np.float32(v) for v in kat["expect_dist"]]
    elif fn == "streaming_top_k":
        ids, sc = O.streaming_top_k(kat["scores"], kat["ids"], kat["k"])
        assert ids == kat["expect_ids"]
        assert sc == [np.float32(v) for v in kat["expect_scores"]]
    elif fn in ("l2", "dot"):
        a = np.full(kat["dim"], kat["a_fill"], np.float32)
        b = np.full(kat["dim"], kat["b_fill"], np.float32)
        got = O.l2(a, b) if fn == "l2" else O.dot(a, b)
        assert abs(got - kat["expect"]) <= kat["tol"]
    elif fn == "cosine":
        assert abs(O.cosine(kat["a"], kat["b"]) - kat["expect"]) <= kat["tol"]
    elif fn == "train":
        data = np.asarray(kat["data"], np.float32)
        init, _ = O.kmeanspp_init(data, kat["n_clusters"], kat["seed"])
        cent, _, res = O.train_lloyd(data, init, kat["max_iterations"])
        assert res["iterations"] == kat["expect_iterations"]
        assert bool(res["converged"]) == kat["expect_converged"]
        assert res["final_error"] < kat["max_final_error"]
        for c in cent:
            assert min(O.l2(c, e) for e in kat["expected_centers"]) < kat["center_tol"]
    elif fn == "ivf_search_count":
        x = np.eye(kat["n_vectors"], 4, dtype=np.float32)
        ivf = O.IVF(x[:1].copy(), x)
        ids, _ = ivf.search(x[0], kat["k"], 1)
        assert len(ids) == kat["expect_count"]
    else:
        raise AssertionError(f"unknown KAT kind {fn}")


def _load(name):
    return np.load(os.path.join(GOLD, f"{name}.npz"))


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_fixture(name):
    n, nf, d, nlist, nq, k, nprobe = make_golden.CASES[name]
    x, q, cents, deleted, keep = make_golden.case_inputs(name)
    g = _load(name)
    ids = np.arange(n, dtype=np.uint32)
    fid = np.arange(n, n + nf, dtype=np.uint32)
    ivf = O.IVF(cents, x[:n], ids)
    assert ivf.assign.tolist() == g["assign"].tolist()
    r = O.hybrid_batch_search(ivf, x[n:], fid, q, k, nprobe, tiers=3)
    assert np.array_equal(r[0], g["plain_ids"]) and np.array_equal(r[2], g["plain_cnt"])
    assert np.array_equal(r[1].view(np.uint32), g["plain_dist"].view(np.uint32))
    r = O.hybrid_batch_search(ivf, x[n:], fid, q, k, nprobe, tiers=3,
                              deleted=O.make_bitmap(n + nf, deleted), filter_bits=O.make_bitmap(n + nf, keep))
    assert np.array_equal(r[0], g["masked_ids"]) and np.array_equal(r[2], g["masked_cnt"])
    assert np.array_equal(r[1].view(np.uint32), g["masked_dist"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["exact", "tc"])
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_cuda_path_reproduces_fixture(name, mode):
    from fabstir_vectordb_b200 import Engine, _lib as L
    n, nf, d, nlist, nq, k, nprobe = make_golden.CASES[name]
    x, q, cents, deleted, keep = make_golden.case_inputs(name)
    g = _load(name)
    eng = Engine(d, k_max=16)
    eng.set_option(L.OPT_SCAN_MODE, L.SCAN_EXACT if mode == "exact" else L.SCAN_TC)
    # Lloyd from the shared initial centroids: bit-identical centroids and iteration count
    res = eng.train(x[:n], nlist, 4, init_centroids=cents)
    assert res["iterations"] == int(g["lloyd_iterations"][0])
    assert np.array_equal(eng.get_centroids().view(np.uint32), g["lloyd_centroids"].view(np.uint32))
    eng.set_centroids(cents)
    lists = eng.ivf_add(x[:n], np.arange(n, dtype=np.uint32), want_lists=True)
    assert lists.tolist() == g["assign"].tolist()
    eng.flat_add(x[n:], np.arange(n, n + nf, dtype=np.uint32))
    ids, dist, cnt = eng.search(q, k, nprobe, tiers=L.TIER_BOTH)
    assert np.array_equal(cnt, g["plain_cnt"]) and np.array_equal(ids, g["plain_ids"])
    assert np.array_equal(dist.view(np.uint32), g["plain_dist"].view(np.uint32))
    eng.set_deleted(deleted, True)
    fb = np.zeros((n + nf + 63) // 64, dtype=np.uint64)
    for i in keep:
        fb[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    ids, dist, cnt = eng.search(q, k, nprobe, tiers=L.TIER_BOTH, filter_bits=fb)
    assert np.array_equal(cnt, g["masked_cnt"]) and np.array_equal(ids, g["masked_ids"])
    assert np.array_equal(dist.view(np.uint32), g["masked_dist"].view(np.uint32))
    eng.close()
