"""Pins the CPU oracle against the reference's own known-answer tests for the hot path
(SURVEY §4 / §8c).  Each test names the reference test it restates."""
import numpy as np

import oracle as O


def test_top_k_selection():
    # tests/core/vector_ops.rs:30-35 test_top_k_selection
    scores = [0.1, 0.9, 0.5, 0.7, 0.3, 0.8]
    assert O.top_k_indices(scores, 3) == [1, 5, 3]
    assert O.top_k_indices(scores, 3, heap=True) == [1, 5, 3]


def test_result_merging():
    # tests/core/vector_ops.rs:38-71 test_result_merging: ids a=0,b=1,c=2
    ids, dist = O.merge_search_results([0, 1, 1, 2], [0.1, 0.3, 0.2, 0.4], 3)
    assert ids == [0, 1, 2]
    assert dist == [np.float32(0.1), np.float32(0.2), np.float32(0.4)]


def test_euclidean_distance_sqrt128():
    # tests/core/vector_ops_advanced.rs:48-58 test_euclidean_distance_simd (scalar arm)
    d = O.l2(np.zeros(128, np.float32), np.ones(128, np.float32))
    assert abs(d - np.sqrt(np.float32(128.0))) < 1e-4


def test_dot_256():
    # tests/core/vector_ops.rs:74-84 test_simd_operations (scalar arm): dot == 256.0 exactly
    assert O.dot(np.ones(256, np.float32), np.ones(256, np.float32)) == 256.0


def test_cosine_batch():
    # tests/core/vector_ops.rs:12-27 test_batch_similarity_calculation
    q = [1.0, 0.0, 0.0]
    assert abs(O.cosine(q, [1.0, 0.0, 0.0]) - 1.0) < 1e-6
    assert abs(O.cosine(q, [0.0, 1.0, 0.0]) - 0.0) < 1e-6
    assert abs(O.cosine(q, [0.707, 0.707, 0.0]) - 0.707) < 0.01
    assert O.cosine(q, [0.0, 0.0, 0.0]) == 0.0  # zero norm -> 0, src/core/vector_ops.rs:44-46


def test_top_k_heap_all_k():
    # tests/core/vector_ops_advanced.rs:86-100 test_top_k_heap_implementation
    scores = [0.9, 0.1, 0.7, 0.3, 0.8, 0.2, 0.6, 0.4, 0.5]
    for k in range(1, len(scores) + 1):
        idx = O.top_k_indices(scores, k, heap=True)
        assert len(idx) == k
        for i in range(1, k):
            assert scores[idx[i - 1]] >= scores[idx[i]]
        assert idx == O.top_k_indices(scores, k)


def test_streaming_top_k():
    # tests/core/vector_ops_advanced.rs:103-124 test_streaming_top_k: ids a..e = 0..4
    ids, sc = O.streaming_top_k([0.5, 0.9, 0.3, 0.7, 0.8], [0, 1, 2, 3, 4], 3)
    assert ids == [1, 4, 3]
    assert sc == [np.float32(0.9), np.float32(0.8), np.float32(0.7)]


def test_euclidean_properties():
    # tests/core/vector_ops.rs:112-136 proptest test_euclidean_distance_properties
    rng = np.random.default_rng(7)
    for _ in range(200):
        n = int(rng.integers(10, 100))
        a = rng.uniform(-100, 100, n).astype(np.float32)
        b = rng.uniform(-100, 100, n).astype(np.float32)
        assert abs(O.l2(a, b) - O.l2(b, a)) < 1e-6 * max(1.0, O.l2(a, b))
        assert O.l2(a, b) >= 0.0
        assert abs(O.l2(a, a)) < 1e-6
        s = O.cosine(a, b)
        assert -1.0 - 1e-6 <= s <= 1.0 + 1e-6
        assert abs(O.cosine(a, a) - 1.0) < 1e-6


def test_l2_is_sequential_f32():
    # the restatement must be a left fold in f32 (src/core/vector_ops.rs:51-57), not pairwise
    rng = np.random.default_rng(3)
    a = rng.standard_normal(384).astype(np.float32)
    b = rng.standard_normal(384).astype(np.float32)
    acc = np.float32(0.0)
    for x, y in zip(a, b):
        t = np.float32(x - y)
        acc = np.float32(acc + np.float32(t * t))
    assert O.l2(a, b) == float(np.sqrt(acc))
    many = O.l2_many(a, np.stack([b] * 11))
    assert all(float(v) == O.l2(a, b) for v in many)


TRAIN_2D = np.array([[0.0, 0.0], [0.1, 0.1], [0.2, -0.1], [5.0, 5.0], [5.1, 4.9], [4.9, 5.1],
                     [-5.0, -5.0], [-4.9, -5.1], [-5.1, -4.9]], dtype=np.float32)


def test_train_simple_2d():
    # tests/ivf/core.rs:69-122 test_train_simple_2d: 9 points / 3 clusters / seed 42 /
    # max_iterations 10 -> iterations == 10 (the n<20 special case), converged, error < 1
    init, picked = O.kmeanspp_init(TRAIN_2D, 3, 42)
    assert init.shape == (3, 2)
    cent, assign, res = O.train_lloyd(TRAIN_2D, init, 10)
    assert res["iterations"] == 10
    assert res["converged"]
    assert res["final_error"] < 1.0
    for c in cent:
        assert min(O.l2(c, e) for e in ([0, 0], [5, 5], [-5, -5])) < 1.0


def test_find_nearest_centroid_ties_lowest_id():
    # src/ivf/core.rs:373-386: strict '<' keeps the first minimum
    cents = np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 5.0]], dtype=np.float32)
    assert O.find_nearest_centroid([1.0, 0.0], cents) == 0
    assert O.assign(np.array([[1.0, 0.0], [0.0, 4.0]], np.float32), cents).tolist() == [0, 2]


def _small_index(seed=0, n=600, d=8, nlist=6):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    cents = x[:nlist].copy()
    return x, cents, O.IVF(cents, x)


def test_search_exact_match_and_sorted():
    # tests/ivf/core.rs:401-413 test_search_exact_match; :339-342 sorted ascending
    x, cents, ivf = _small_index()
    ids, dist = ivf.search(x[17], 5, 6)
    assert ids[0] == 17 and dist[0] < 1e-6
    assert all(dist[i] <= dist[i + 1] for i in range(len(dist) - 1))


def test_search_more_k_than_vectors():
    # tests/ivf/core.rs:386-398 test_search_more_k_than_vectors: 3 vectors, k=10 -> 3 results
    cents = np.array([[0.0, 0.0], [10.0, 10.0]], np.float32)
    x = np.array([[0.1, 0.0], [0.0, 0.2], [9.0, 9.5]], np.float32)
    ivf = O.IVF(cents, x)
    ids, dist = ivf.search([0.0, 0.0], 10, 2)
    assert len(ids) == 3


def test_search_multi_probe_monotone():
    # tests/ivf/core.rs:346-383 test_search_multi_probe: more probes never lose results
    x, cents, ivf = _small_index(1)
    q = x[3] + 0.05
    d1 = ivf.search(q, 10, 1)[1]
    d3 = ivf.search(q, 10, 3)[1]
    d6 = ivf.search(q, 10, 6)[1]
    assert len(d3) >= len(d1)
    assert d3[0] <= d1[0] and d6[0] <= d3[0]
    # all lists probed == exact scan
    fi, fd = O.flat_search(x, None, q, 10)
    i6, _ = ivf.search(q, 10, 6)
    assert i6.tolist() == fi.tolist() and d6.tolist() == fd.tolist()


def test_deleted_rows_are_skipped():
    # tests/unit/ivf_deletion_tests.rs:102-128: soft-deleted ids never come back
    x, cents, ivf = _small_index(2)
    q = x[5]
    ids, _ = ivf.search(q, 5, 6)
    dele = O.make_bitmap(len(x), ids[:2])
    ids2, _ = ivf.search(q, 5, 6, deleted=dele)
    assert not set(ids[:2].tolist()) & set(ids2.tolist())
    assert ids2[0] == ids[2]


def test_faithful_cost_mode_same_answer():
    x, cents, ivf = _small_index(4)
    for i in range(5):
        a = ivf.search(x[i] + 0.01, 7, 3)
        b = ivf.search(x[i] + 0.01, 7, 3, faithful=True)
        assert a[0].tolist() == b[0].tolist() and a[1].tolist() == b[1].tolist()


def test_hybrid_merge_recent_first_no_dedup():
    # src/hybrid/core.rs:456-483: recent results first on ties; the same vector in both tiers
    # is returned twice (no dedup, Appendix A.5)
    x, cents, ivf = _small_index(5)
    flat = x[:50].copy()
    flat_ids = np.arange(50, dtype=np.uint32) + 10_000
    ids, dist, cnt = O.hybrid_batch_search(ivf, flat, flat_ids, x[:4], 4, 6)
    for i in range(4):
        assert cnt[i] == 4
        assert ids[i, 0] == 10_000 + i and ids[i, 1] == i
        assert dist[i, 0] == dist[i, 1] == 0.0


def test_postfilter_is_subset_of_prefilter():
    # src/hybrid/core.rs:513-549 (3x oversample post-filter) vs the bitmap pre-filter
    x, cents, ivf = _small_index(6, n=2000)
    match = O.make_bitmap(len(x), np.arange(0, len(x), 10))
    for i in range(8):
        q = x[i] + 0.02
        post_ids, _ = O.hybrid_search_postfilter(ivf, None, None, q, 5, 6, match)
        pre_ids, pre_d, pre_c = O.hybrid_batch_search(ivf, None, None, q[None, :], 5, 6,
                                                      filter_bits=match)
        pre = pre_ids[0, :pre_c[0]].tolist()
        assert len(post_ids) <= 5 and pre_c[0] == 5
        assert post_ids.tolist() == pre[:len(post_ids)]
        assert all(int(j) % 10 == 0 for j in pre)


def test_recall_definition():
    # src/ivf/operations.rs:355-371
    found = np.array([[1, 2, 3, 9]], np.uint32)
    truth = np.array([[1, 2, 3, 4]], np.uint32)
    assert O.recall(found, [4], truth, [4], 4) == 0.75


def test_update_centroids_empty_cluster_keeps_old():
    # src/ivf/core.rs:410-415
    data = np.array([[1.0, 1.0], [3.0, 3.0]], np.float32)
    cents = np.array([[0.0, 0.0], [9.0, 9.0]], np.float32)
    out = O.update_centroids(data, [0, 0], cents)
    assert out.tolist() == [[2.0, 2.0], [9.0, 9.0]]


def test_stdrng_stream_is_deterministic():
    a = O.stdrng_stream(42, 130)
    b = O.stdrng_stream(42, 130)
    assert a.tolist() == b.tolist() and a.tolist() != O.stdrng_stream(43, 130).tolist()


def test_retrain_restatement_properties():
    """retrain (src/ivf/operations.rs:148-193): same vectors and ids afterwards, every vector in the
    list of its nearest NEW centroid (reinsert -> insert, src/ivf/core.rs:431-455), centroid count
    = the new config's; training order is list-major (stable)."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((400, 8)).astype(np.float32)
    cents = x[:6].copy()
    ivf = O.IVF(cents, x, np.arange(400, dtype=np.uint32) + 1000)
    init = x[rng.choice(400, 9, replace=False)].copy()
    new_ivf, order, res = O.retrain_lloyd(ivf, init, 7)
    assert new_ivf.nlist == 9 and new_ivf.n == 400
    assert sorted(new_ivf.ids.tolist()) == sorted(ivf.ids.tolist())
    assert (np.diff(ivf.assign[order].astype(np.int64)) >= 0).all()
    assert new_ivf.assign.tolist() == O.assign(new_ivf.rows, new_ivf.centroids).tolist()
    assert 1 <= res["iterations"] <= 7
    # the training data is exactly the stored vectors, reordered
    assert np.array_equal(new_ivf.rows, x[order])


def test_flat_search_metric_is_batch_cosine_plus_top_k_indices():
    """The metric search of the oracle = batch_cosine_similarity (tests/core/vector_ops.rs:8-27: 1.0, 0.0,
    ~0.707) followed by top_k_indices (:29-35: a stable descending sort); zero vectors score 0
    (src/core/vector_ops.rs:44-46); ties keep the input order."""
    q = np.array([[1.0, 0.0, 0.0]], dtype=np.float32)
    rows = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.707, 0.707, 0.0], [0.0, 0.0, 0.0], [2.0, 0.0, 0.0]],
                    dtype=np.float32)
    ids = np.arange(5, dtype=np.uint32)
    i, s, c = O.flat_search_metric(rows, ids, q, 5, O.COSINE)
    assert c.tolist() == [5]
    assert i[0].tolist() == [0, 4, 2, 1, 3]            # 1.0, 1.0 (tie: input order), 0.707, 0.0, 0.0 (tie)
    assert abs(s[0, 0] - 1.0) < 1e-6 and abs(s[0, 2] - 0.707) < 0.01 and s[0, 3] == 0.0 and s[0, 4] == 0.0
    # scores are the scalar kernels' own values, the order is top_k_indices' over them
    rng = np.random.default_rng(3)
    x = rng.standard_normal((200, 16)).astype(np.float32)
    qq = rng.standard_normal((3, 16)).astype(np.float32)
    for metric, fn in ((O.COSINE, O.cosine), (O.DOT, O.dot)):
        i, s, c = O.flat_search_metric(x, np.arange(200, dtype=np.uint32), qq, 7, metric)
        for j in range(3):
            scores = np.array([fn(qq[j], x[r]) for r in range(200)], dtype=np.float32)
            assert i[j].tolist() == O.top_k_indices(scores, 7)
            assert s[j].view(np.uint32).tolist() == scores[i[j]].view(np.uint32).tolist()
    # deleted rows and the filter bitmap are skipped before scoring
    dele = O.make_bitmap(200, [int(i[0, 0])])
    i2, _, _ = O.flat_search_metric(x, np.arange(200, dtype=np.uint32), qq[:1], 7, O.DOT, deleted=dele)
    assert int(i[0, 0]) not in i2[0].tolist()
