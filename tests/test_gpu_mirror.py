"""GPU tests of the host mirror of the reference's index API (fabstir_vectordb_b200/index.py:
IVFIndex / HNSWIndex / HybridIndex) and of the 3x post-filter entry of the C ABI, against the CPU
oracle.  The mirror keeps only what the reference keeps host-side (VectorId <-> row id, timestamps,
the deleted set); every number comes from the engine."""
import time

import numpy as np
import pytest

import oracle as O
from fabstir_vectordb_b200 import (DuplicateVector, Engine, FvdbError, HNSWIndex, HybridConfig, HybridIndex,
                                   HybridSearchConfig, IVFConfig, IVFIndex, InvalidParameter, NanInput,
                                   VectorNotFound, _lib as L, synth)

pytestmark = pytest.mark.gpu

D = 384


def _rows(n, seed, n_comp=32, sigma=0.6):
    return synth.rows(0, n, D, n_comp, sigma, seed)


def _queries(nq, n, seed, n_comp=32, sigma=0.6):
    return synth.queries(0, nq, D, n, n_comp, sigma, seed, synth.default_qnoise(D, sigma), seed + 1)


def _pairs(results):
    return [(r.vector_id, np.float32(r.distance).view(np.uint32)) for r in results]


def test_hnsw_mirror_search_is_the_exact_scan():
    """HNSWIndex::search (src/hnsw/core.rs:398-467) served by the exhaustive scan: ids and distance
    bits equal the oracle's flat scan; deleted nodes are skipped (:452-459); empty index -> []."""
    n = 5000
    x = _rows(n, 3)
    idx = HNSWIndex(k_max=32)
    assert idx.search(x[0], 5) == []
    vids = [f"v{i}" for i in range(n)]
    idx.batch_insert(vids, x)
    assert idx.node_count() == n
    dele = [7, 11, 4093]
    for r in dele:
        idx.mark_deleted(vids[r])
    with pytest.raises(VectorNotFound):
        idx.mark_deleted("nope")
    q = _queries(12, n, 3)
    rows = np.arange(n, dtype=np.uint32)
    want = O.hybrid_batch_search(None, x, rows, q, 10, 0, tiers=1, deleted=O.make_bitmap(n, np.asarray(dele, np.uint32)))
    for i in range(q.shape[0]):
        got = idx.search(q[i], 10, ef=50)
        assert [g.vector_id for g in got] == [vids[int(r)] for r in want[0][i, :want[2][i]]]
        assert [np.float32(g.distance).view(np.uint32) for g in got] == want[1][i, :want[2][i]].view(np.uint32).tolist()


def _hybrid(n_recent, n_hist, nlist, seed):
    n = n_recent + n_hist
    x = _rows(n, seed)
    cfg = HybridConfig(ivf_config=IVFConfig(n_clusters=nlist, n_probe=4, train_size=1000, max_iterations=5),
                       auto_migrate=False)
    idx = HybridIndex(cfg, k_max=64)
    init = x[:: n // nlist][:nlist].copy()
    idx.initialize(x[:2000], init_centroids=init)
    now = time.time()
    vids = [("id", i) for i in range(n)]
    # the first n_hist vectors are older than recent_threshold -> IVF tier, the rest -> recent tier
    ts = [now - 30 * 24 * 3600.0] * n_hist + [now] * n_recent
    idx.batch_insert_with_timestamps(vids, x, ts)
    assert idx.recent_count() == n_recent and idx.historical_count() == n_hist
    cents = idx.engine().get_centroids()
    ivf = O.IVF(cents, x[:n_hist], np.arange(n_hist, dtype=np.uint32))
    return idx, ivf, x, vids


def test_hybrid_mirror_search_matches_the_oracle():
    """HybridIndex::search / search_with_config (src/hybrid/core.rs:419-486): recent tier ∪ IVF tier,
    stable sort with the recent tier first on ties, truncate(k); default ivf_n_probe = 10."""
    n_recent, n_hist, nlist = 1500, 6000, 24
    idx, ivf, x, vids = _hybrid(n_recent, n_hist, nlist, 9)
    q = _queries(20, n_recent + n_hist, 9)
    fr, fi = x[n_hist:], np.arange(n_hist, n_hist + n_recent, dtype=np.uint32)
    want = O.hybrid_batch_search(ivf, fr, fi, q, 10, 10, tiers=3)
    got = idx.batch_search(q, 10)
    for i in range(q.shape[0]):
        c = int(want[2][i])
        assert [g.vector_id for g in got[i]] == [vids[int(r)] for r in want[0][i, :c]]
        assert [np.float32(g.distance).view(np.uint32) for g in got[i]] == want[1][i, :c].view(np.uint32).tolist()
    # one tier at a time, explicit n_probe (SearchConfig, :173-196)
    w_h = O.hybrid_batch_search(ivf, fr, fi, q[:4], 5, 3, tiers=2)
    g_h = idx.batch_search_with_config(q[:4], HybridSearchConfig(k=5, ivf_n_probe=3, search_recent=False))
    w_r = O.hybrid_batch_search(ivf, fr, fi, q[:4], 5, 3, tiers=1)
    g_r = idx.batch_search_with_config(q[:4], HybridSearchConfig(k=5, search_historical=False))
    for i in range(4):
        assert [g.vector_id for g in g_h[i]] == [vids[int(r)] for r in w_h[0][i, :w_h[2][i]]]
        assert [g.vector_id for g in g_r[i]] == [vids[int(r)] for r in w_r[0][i, :w_r[2][i]]]
    # single-query entry == row of the batch
    assert _pairs(idx.search(q[0], 10)) == _pairs(got[0])


def test_search_with_filter_is_the_reference_post_filter():
    """HybridIndex::search_with_filter (src/hybrid/core.rs:513-549) through fvdb_search_postfilter ==
    the oracle's fo_hybrid_search_postfilter: search(3k), keep matches, truncate(k) — so it can return
    fewer than k although more matches exist, while the in-kernel pre-filter cannot."""
    n_recent, n_hist, nlist = 1200, 5000, 16
    idx, ivf, x, vids = _hybrid(n_recent, n_hist, nlist, 17)
    n = n_recent + n_hist
    # metadata: 10 % of the rows are "tech"; every 13th row has no metadata at all (dropped, :536-541)
    meta = {str(vids[i]): {"cat": "tech" if i % 10 == 0 else "other"} for i in range(n) if i % 13}
    flt = lambda md: md["cat"] == "tech"                                      # noqa: E731
    match_rows = np.asarray([i for i in range(n) if i % 13 and i % 10 == 0], dtype=np.uint32)
    match_bits = O.make_bitmap(n, match_rows)
    assert np.array_equal(idx.filter_bitmap(flt, meta)[:match_bits.size], match_bits)
    q = _queries(16, n, 17)
    fr, fi = x[n_hist:], np.arange(n_hist, n, dtype=np.uint32)
    k = 10
    short = 0
    for i in range(q.shape[0]):
        w_ids, w_dist = O.hybrid_search_postfilter(ivf, fr, fi, q[i], k, 10, match_bits, tiers=3)
        got = idx.search_with_filter(q[i], k, flt, meta)
        assert [g.vector_id for g in got] == [vids[int(r)] for r in w_ids]
        assert [np.float32(g.distance).view(np.uint32) for g in got] == w_dist.view(np.uint32).tolist()
        pre = idx.search_with_prefilter(q[i], k, flt, meta)
        assert len(pre) == k                                   # >= k matching rows are reachable
        assert {g.vector_id for g in got} <= {p.vector_id for p in pre} | {g.vector_id for g in got}
        short += len(got) < k
    assert short > 0, "at 10 % selectivity the 3x post-filter must come up short for some query"
    assert _pairs(idx.search_with_filter(q[0], k, None, meta)) == _pairs(idx.search(q[0], k))
    with pytest.raises(InvalidParameter):
        idx.search_with_filter(q[0], 30, flt, meta)            # 3k = 90 > k_max = 64
    # the batch entry of the ABI: same rows for all queries in one call, tombstones respected
    eng = idx.engine()
    dele = match_rows[:40]
    eng.set_deleted(dele, True)
    g_ids, g_dist, g_cnt = eng.search_postfilter(q, k, 10, match_bits, tiers=L.TIER_BOTH)
    for i in range(q.shape[0]):
        w_ids, w_dist = O.hybrid_search_postfilter(ivf, fr, fi, q[i], k, 10, match_bits, tiers=3,
                                                   deleted=O.make_bitmap(n, dele))
        assert g_cnt[i] == len(w_ids)
        assert g_ids[i, :g_cnt[i]].tolist() == w_ids.tolist()
        assert g_dist[i, :g_cnt[i]].view(np.uint32).tolist() == w_dist.view(np.uint32).tolist()
    with pytest.raises(FvdbError):
        eng.search_postfilter(q, 22, 10, match_bits)           # FVDB_ERR_K_TOO_LARGE: 66 > 64


def test_vacuum_releases_the_vector_ids():
    """After delete + vacuum the reference has removed the entry physically (src/ivf/operations.rs:625-645):
    the same VectorId can be inserted again, counts drop, and a deleted id is unknown."""
    x = _rows(600, 23)
    # IVF mirror
    ivf = IVFIndex(IVFConfig(n_clusters=8, n_probe=8, train_size=600, max_iterations=3), k_max=16)
    ivf.train(x, init_centroids=x[:8].copy())
    ivf.batch_insert(list(range(500)), x[:500])
    ivf.mark_deleted(42)
    assert ivf.vacuum() == 1
    assert ivf.total_vectors() == 499
    with pytest.raises(VectorNotFound):
        ivf.mark_deleted(42)
    ivf.insert(42, x[550])                                      # re-insert under the same id, new vector
    assert ivf.total_vectors() == 500
    res = ivf.search_with_config(x[550], 1, 8)
    assert res[0].vector_id == 42 and res[0].distance == 0.0
    # HNSW mirror
    hn = HNSWIndex(k_max=16)
    hn.batch_insert([f"a{i}" for i in range(100)], x[:100])
    hn.mark_deleted("a5")
    assert hn.vacuum() == 1
    assert hn.node_count() == 99
    with pytest.raises(VectorNotFound):
        hn.mark_deleted("a5")
    hn.insert("a5", x[560])
    assert hn.node_count() == 100
    assert hn.search(x[560], 1)[0].vector_id == "a5"
    # Hybrid mirror
    hy = HybridIndex(HybridConfig(ivf_config=IVFConfig(n_clusters=4, n_probe=4, train_size=100, max_iterations=3),
                                  auto_migrate=False), k_max=16)
    hy.initialize(x[:100], init_centroids=x[:4].copy())
    now = time.time()
    hy.batch_insert_with_timestamps(list(range(200)), x[:200], [now - 1e7] * 100 + [now] * 100)
    hy.delete(3)
    hy.delete(150)
    assert hy.vacuum() == 2
    assert 3 not in hy.timestamps and 150 not in hy.timestamps
    hy.insert_with_timestamp(3, x[570], now)
    hy.insert_with_timestamp(150, x[571], now - 1e7)
    assert hy.recent_count() == 100 and hy.historical_count() == 100
    assert hy.search(x[570], 1)[0].vector_id == 3
    assert hy.search(x[571], 1)[0].vector_id == 150


def test_failed_batch_insert_leaves_nothing_behind():
    x = _rows(300, 29)
    hy = HybridIndex(HybridConfig(ivf_config=IVFConfig(n_clusters=4, n_probe=4, train_size=100, max_iterations=3),
                                  auto_migrate=False), k_max=16)
    hy.initialize(x[:100], init_centroids=x[:4].copy())
    now = time.time()
    hy.batch_insert_with_timestamps(["a", "b"], x[:2], [now, now - 1e7])
    with pytest.raises(DuplicateVector):                        # duplicate INSIDE the batch
        hy.batch_insert_with_timestamps(["c", "d", "c"], x[2:5], [now] * 3)
    with pytest.raises(DuplicateVector):                        # duplicate against the index
        hy.batch_insert_with_timestamps(["e", "a"], x[5:7], [now] * 2)
    bad = x[7:11].copy()
    bad[3, 5] = np.nan
    with pytest.raises(NanInput):                               # NaN in the IVF half of a mixed batch
        hy.batch_insert_with_timestamps(["f", "g", "h", "i"], bad, [now, now, now - 1e7, now - 1e7])
    assert hy.recent_count() == 1 and hy.historical_count() == 1
    # every id of the failed batches is free again
    hy.batch_insert_with_timestamps(["c", "d", "e", "f", "g", "h", "i"], x[20:27], [now] * 7)
    assert hy.recent_count() == 8
    st = hy.engine().stats()
    assert st.flat_rows == 8 and st.ivf_rows == 1
    hn = HNSWIndex(k_max=16)
    hn.batch_insert(["p"], x[:1])
    with pytest.raises(NanInput):
        hn.batch_insert(["q", "r"], bad[2:4])
    hn.batch_insert(["q", "r"], x[30:32])
    assert hn.node_count() == 3


def test_rejected_train_leaves_the_index_as_it_was():
    """IVFIndex::train validates before it touches anything (src/ivf/core.rs:242-262): a rejected call on
    a trained, populated handle must leave centroids, lists and the trained flag alone."""
    x = _rows(3000, 31)
    eng = Engine(D, k_max=16)
    cents = x[:16].copy()
    eng.set_centroids(cents)
    eng.ivf_add(x, np.arange(3000, dtype=np.uint32))
    q = _queries(8, 3000, 31)
    before = eng.search(q, 5, 4, tiers=L.TIER_HISTORICAL)
    bad_init = x[:32].copy()
    bad_init[7, 7] = np.nan
    with pytest.raises(NanInput):
        eng.train(x, 32, 3, init_centroids=bad_init)
    st = eng.stats()
    assert st.trained == 1 and st.nlist == 16 and st.ivf_rows == 3000
    assert np.array_equal(eng.get_centroids().view(np.uint32), cents.view(np.uint32))
    after = eng.search(q, 5, 4, tiers=L.TIER_HISTORICAL)
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[1].view(np.uint32), after[1].view(np.uint32))
