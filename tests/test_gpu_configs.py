"""Parity at the FULL sizes of BASELINE.json's configs, run by `pytest -m gpu`:

  cfg 2  1M x 384 IVF, nlist 1024, nprobe 32, k 10: 128 queries of a 1024-query batch against the oracle
         (bit-exact), the whole batch through size-independent properties
  cfg 3  k-means at 1M x 384, nlist 4096: one full Lloyd iteration bit-exact against the oracle's
         update_centroids given sample-verified assignments; plus the reduced end-to-end run of SURVEY §8d
         (100K rows, nlist 512, 20 iterations, shared init) bit-exact against the oracle's Lloyd loop
  cfg 4  300K recent + 700K IVF, 10 % filter bitmap, 1 % tombstones: 64 queries against the oracle

The oracle costs ~0.5 us per 384-d distance, so whatever it checks at these sizes is a bounded sample;
the assignment the oracle's IVF is built from is the engine's, verified against the oracle on 20 K rows.
"""
import os
import sys

import numpy as np
import pytest

import oracle as O
from fabstir_vectordb_b200 import Engine, _lib as L, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIM, SEED, SIGMA = 384, 1234, 1.0


def _torch():
    import torch
    torch.cuda.set_device(0)
    return torch


def _gen_rows_host(torch, lib, r0, n, n_comp):
    buf = torch.empty((n, DIM), dtype=torch.float32, device="cuda")
    assert lib.fvdb_synth_rows_device(buf.data_ptr(), r0, n, DIM, n_comp, SIGMA, SEED,
                                      torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    return buf


def _same(got, want):
    g_ids, g_dist, g_cnt = got
    w_ids, w_dist, w_cnt = want
    assert g_cnt.tolist() == w_cnt.tolist()
    for i in range(len(w_cnt)):
        c = int(w_cnt[i])
        assert g_ids[i, :c].tolist() == w_ids[i, :c].tolist(), f"query {i}"
        assert g_dist[i, :c].view(np.uint32).tolist() == w_dist[i, :c].view(np.uint32).tolist(), f"query {i}"


@pytest.fixture(scope="module")
def bench_mod():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_cfg2_1m_ivf_search_parity(bench_mod):
    bench = bench_mod
    torch = _torch()
    lib = L.load()
    eng = Engine(DIM, k_max=16)
    n_total, nlist, n_comp = bench.build_index(torch, eng, 0, 1, lambda m: None)
    assert (n_total, nlist) == (1_000_000, 1024)
    ivf, x = bench.host_index_from_device(torch, lib, eng, n_total, n_comp)
    # the oracle's lists come from the engine's assignment: verify it on a 20 K-row sample
    rows = np.random.default_rng(1).choice(n_total, 20_000, replace=False)
    assert np.array_equal(O.assign(x[rows], ivf.centroids), ivf.assign[rows])
    q_dev = bench.make_queries(torch, lib, bench.NQ_PER_GPU, n_total, n_comp, 0)
    q = q_dev.cpu().numpy()
    got = eng.search(q, bench.K, bench.NPROBE, tiers=L.TIER_HISTORICAL)
    assert eng.stats().last_fallback_queries == 0
    want = O.hybrid_batch_search(ivf, None, None, q[:128], bench.K, bench.NPROBE, tiers=2)
    _same(tuple(a[:128] for a in got), want)
    # the whole 1024-query batch: full counts, ascending distances, no repeated row, every distance
    # the exact fp32 distance of the row it names (recomputed here in the reference's order on 64 rows)
    ids, dist, cnt = got
    assert (cnt == bench.K).all()
    assert (np.diff(dist, axis=1) >= 0).all()
    assert all(len(set(r.tolist())) == bench.K for r in ids)
    for qi in range(0, 1024, 16):
        r = int(ids[qi, 3])
        assert np.float32(O.l2(q[qi], x[r])).view(np.uint32) == dist[qi, 3].view(np.uint32)
    # the exact CUDA-core mode returns the same bits for the whole batch
    eng.set_option(L.OPT_SCAN_MODE, L.SCAN_EXACT)
    ex = eng.search(q, bench.K, bench.NPROBE, tiers=L.TIER_HISTORICAL)
    assert np.array_equal(ex[0], ids) and np.array_equal(ex[1].view(np.uint32), dist.view(np.uint32))
    eng.close()


def test_cfg3_kmeans_full_size_iteration_parity():
    """1M x 384, nlist 4096: fvdb_ivf_train with max_iterations = 1 is assign (against the shared init)
    + update_centroids.  The assignment is checked against the oracle on 20 K rows, the update — all
    4096 x 384 centroid words — bit for bit given that assignment (src/ivf/core.rs:388-417)."""
    torch = _torch()
    lib = L.load()
    n, nlist = 1_000_000, 4096
    n_comp = 4 * nlist
    data = _gen_rows_host(torch, lib, 0, n, n_comp)
    init_d = data[torch.arange(nlist, device="cuda") * (n // nlist)].contiguous()
    init = init_d.cpu().numpy()
    eng = Engine(DIM, k_max=16)
    res = eng.train_device(data.data_ptr(), n, nlist, 1, init_d.data_ptr(), SEED)
    assert res["iterations"] == 1
    cents1 = eng.get_centroids()
    x = data.cpu().numpy()
    e2 = Engine(DIM, k_max=16)
    e2.set_centroids(init)
    a = e2.assign(x)
    rows = np.random.default_rng(2).choice(n, 20_000, replace=False)
    assert np.array_equal(O.assign(x[rows], init), a[rows])
    want = O.update_centroids(x, a, init)
    assert np.array_equal(cents1.view(np.uint32), want.view(np.uint32))
    # errors of the TrainResult: initial (all points on centroid 0, :280-282) and final, on a prefix the
    # oracle can afford — the engine's f32 left fold over the same prefix
    m = 50_000
    e3 = Engine(DIM, k_max=16)
    r3 = e3.train(x[:m], 64, 1, init_centroids=init[:64])
    o_c, _, o_r = O.train_lloyd(x[:m], init[:64], 1)
    assert np.float32(r3["initial_error"]).view(np.uint32) == np.float32(o_r["initial_error"]).view(np.uint32)
    assert np.float32(r3["final_error"]).view(np.uint32) == np.float32(o_r["final_error"]).view(np.uint32)
    for e in (eng, e2, e3):
        e.close()


def test_cfg3_kmeans_reduced_end_to_end_parity():
    """N = 100K, nlist = 512, 20 iterations, shared init — centroids, iteration count,
    convergence flag and errors equal the oracle's Lloyd loop bit for bit."""
    torch = _torch()
    lib = L.load()
    n, nlist, iters = 100_000, 512, 20
    x = _gen_rows_host(torch, lib, 0, n, 4 * nlist).cpu().numpy()
    init = x[:: n // nlist][:nlist].copy()
    eng = Engine(DIM, k_max=16)
    res = eng.train(x, nlist, iters, init_centroids=init)
    o_c, _, o_r = O.train_lloyd(x, init, iters)
    assert res["iterations"] == o_r["iterations"] and res["converged"] == o_r["converged"]
    assert np.array_equal(eng.get_centroids().view(np.uint32), o_c.view(np.uint32))
    assert np.float32(res["final_error"]).view(np.uint32) == np.float32(o_r["final_error"]).view(np.uint32)
    eng.close()


def test_cfg4_filtered_hybrid_parity(bench_mod):
    bench = bench_mod
    torch = _torch()
    lib = L.load()
    n_total, n_recent, nlist, nq, k, nprobe = 1_000_000, 300_000, 1024, 1024, 10, 32
    n_ivf = n_total - n_recent
    n_comp = 4 * nlist
    eng = Engine(DIM, k_max=16)
    n_train = 64 * nlist
    train = torch.empty((n_train, DIM), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.fvdb_synth_rows_strided_device(train.data_ptr(), 0, n_train, DIM, n_comp, SIGMA, SEED, 64,
                                              max(1, n_total // n_train), stream) == 0
    torch.cuda.synchronize()
    blk = torch.arange(nlist, device="cuda")
    init = train[blk * 64 + (blk // max(1, nlist // 16)) % 64].contiguous()
    eng.train_device(train.data_ptr(), n_train, nlist, 8, init.data_ptr(), SEED)
    CH = 1 << 18
    x = np.empty((n_total, DIM), dtype=np.float32)
    for r0 in range(0, n_total, CH):
        n = min(CH, n_total - r0)
        buf = _gen_rows_host(torch, lib, r0, n, n_comp)
        ids = torch.arange(r0, r0 + n, dtype=torch.int32, device="cuda")
        a = max(0, min(n, n_ivf - r0))
        if a > 0:
            eng.ivf_add_device(buf.data_ptr(), ids.data_ptr(), a)
        if a < n:
            eng.flat_add_device(buf[a:].data_ptr(), ids[a:].data_ptr(), n - a)
        x[r0:r0 + n] = buf.cpu().numpy()
    fbits = synth.filter_bitmap((n_total + 63) // 64 * 64, 10, 99)            # 10 % of the rows pass
    dele = np.arange(7, n_total, 100, dtype=np.uint32)                          # 1 % tombstones
    eng.set_deleted(dele, True)
    q = bench.make_queries(torch, lib, nq, n_total, n_comp, 0).cpu().numpy()
    got = eng.search(q, k, nprobe, tiers=L.TIER_BOTH, filter_bits=fbits)
    assert eng.stats().last_fallback_queries == 0
    cents = eng.get_centroids()
    assign = eng.assign(x[:n_ivf])
    rows = np.random.default_rng(3).choice(n_ivf, 20_000, replace=False)
    assert np.array_equal(O.assign(x[rows], cents), assign[rows])
    ivf = O.IVF(cents, x[:n_ivf], np.arange(n_ivf, dtype=np.uint32), assign_=assign)
    ns = 64
    want = O.hybrid_batch_search(ivf, x[n_ivf:], np.arange(n_ivf, n_total, dtype=np.uint32), q[:ns], k, nprobe,
                                 tiers=3, deleted=O.make_bitmap(n_total, dele), filter_bits=fbits)
    _same(tuple(a[:ns] for a in got), want)
    # whole batch: every returned row passes the filter and is not tombstoned
    ids, dist, cnt = got
    dead = set(dele.tolist())
    for qi in range(nq):
        for r in ids[qi, :cnt[qi]].tolist():
            assert (int(fbits[r >> 6]) >> (r & 63)) & 1 and r not in dead
    assert (np.diff(dist, axis=1)[cnt[:, None] > np.arange(1, k)[None, :]] >= 0).all()
    eng.close()
