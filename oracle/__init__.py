"""ctypes binding of oracle/libfvdb_oracle.so — the CPU restatement of the reference path.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Each wrapper cites the reference lines its C function follows (see fvdb_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfvdb_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fvdb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def _declare(L):
    L.fo_l2.restype = C.c_float
    L.fo_l2.argtypes = [_f32p, _f32p, C.c_size_t]
    L.fo_dot.restype = C.c_float
    L.fo_dot.argtypes = [_f32p, _f32p, C.c_size_t]
    L.fo_cosine.restype = C.c_float
    L.fo_cosine.argtypes = [_f32p, _f32p, C.c_size_t]
    L.fo_l2_many.restype = None
    L.fo_l2_many.argtypes = [_f32p, _f32p, C.c_size_t, C.c_size_t, _f32p]
    for name in ("fo_top_k_indices", "fo_top_k_indices_heap"):
        f = getattr(L, name)
        f.restype = C.c_size_t
        f.argtypes = [_f32p, C.c_size_t, C.c_size_t, _u32p]
    L.fo_streaming_top_k.restype = C.c_size_t
    L.fo_streaming_top_k.argtypes = [_f32p, _u32p, C.c_size_t, C.c_size_t, _u32p, _f32p]
    L.fo_merge_search_results.restype = C.c_size_t
    L.fo_merge_search_results.argtypes = [_u32p, _f32p, C.c_size_t, C.c_size_t, _u32p, _f32p]
    L.fo_find_nearest_centroid.restype = C.c_uint32
    L.fo_find_nearest_centroid.argtypes = [_f32p, _f32p, C.c_size_t, C.c_size_t]
    L.fo_assign.restype = None
    L.fo_assign.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, C.c_size_t, _u32p]
    L.fo_ivf_build.restype = C.c_void_p
    L.fo_ivf_build.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p, _u32p, C.c_size_t, _u32p]
    L.fo_ivf_build_assigned.restype = C.c_void_p
    L.fo_ivf_build_assigned.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p, _u32p, _u32p,
                                        C.c_size_t]
    L.fo_ivf_free.restype = None
    L.fo_ivf_free.argtypes = [C.c_void_p]
    L.fo_ivf_list_len.restype = C.c_size_t
    L.fo_ivf_list_len.argtypes = [C.c_void_p, C.c_size_t]
    L.fo_ivf_coarse.restype = C.c_size_t
    L.fo_ivf_coarse.argtypes = [C.c_void_p, _f32p, C.c_size_t, _u32p, _f32p]
    L.fo_ivf_search.restype = C.c_size_t
    L.fo_ivf_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, _u64p, C.c_uint64,
                                _u64p, C.c_uint64, _u32p, _f32p]
    L.fo_ivf_search_faithful.restype = C.c_size_t
    L.fo_ivf_search_faithful.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, _u64p,
                                         C.c_uint64, _u32p, _f32p]
    L.fo_flat_search.restype = C.c_size_t
    L.fo_flat_search.argtypes = [_f32p, _u32p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, _u64p,
                                 C.c_uint64, _u64p, C.c_uint64, _u32p, _f32p]
    L.fo_hybrid_search.restype = C.c_size_t
    L.fo_hybrid_search.argtypes = [C.c_void_p, _f32p, _u32p, C.c_size_t, C.c_size_t, _f32p,
                                   C.c_size_t, C.c_size_t, C.c_uint, _u64p, C.c_uint64, _u64p,
                                   C.c_uint64, _u32p, _f32p]
    L.fo_hybrid_search_postfilter.restype = C.c_size_t
    L.fo_hybrid_search_postfilter.argtypes = L.fo_hybrid_search.argtypes
    L.fo_flat_search_metric.restype = C.c_size_t
    L.fo_flat_search_metric.argtypes = [_f32p, _u32p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, C.c_int, _u64p,
                                        C.c_uint64, _u64p, C.c_uint64, _u32p, _f32p]
    L.fo_hybrid_batch_search.restype = None
    L.fo_hybrid_batch_search.argtypes = [C.c_void_p, _f32p, _u32p, C.c_size_t, C.c_size_t, _f32p,
                                         C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, _u64p,
                                         C.c_uint64, _u64p, C.c_uint64, C.c_int, C.c_int, _u32p,
                                         _f32p, _u32p]
    L.fo_num_threads.restype = C.c_int
    L.fo_recall.restype = C.c_double
    L.fo_recall.argtypes = [_u32p, _u32p, _u32p, _u32p, C.c_size_t, C.c_size_t]
    L.fo_compute_error.restype = C.c_float
    L.fo_compute_error.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p, _u32p]
    L.fo_update_centroids.restype = None
    L.fo_update_centroids.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_size_t, _u32p, _f32p]
    L.fo_train_lloyd.restype = None
    L.fo_train_lloyd.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, _f32p,
                                 _u32p, _u32p, _u32p, _f32p, _f32p]
    L.fo_stdrng_stream.restype = None
    L.fo_stdrng_stream.argtypes = [C.c_uint64, _u32p, C.c_size_t]
    L.fo_kmeanspp_init.restype = C.c_size_t
    L.fo_kmeanspp_init.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, _f32p,
                                   _u32p]


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(_u32p)


def _bits(a):
    if a is None:
        return None, None, 0
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(_u64p), a.size * 64


def make_bitmap(nbits: int, set_ids) -> np.ndarray:
    """u64 bitmap over row ids with the given ids set."""
    words = np.zeros((nbits + 63) // 64, dtype=np.uint64)
    ids = np.asarray(set_ids, dtype=np.uint64)
    if ids.size:
        np.bitwise_or.at(words, (ids >> np.uint64(6)).astype(np.int64),
                         np.uint64(1) << (ids & np.uint64(63)))
    return words


# ---- scalar kernels (src/core/vector_ops.rs:35-57) ---------------------------------------
def l2(a, b) -> float:
    a, pa = _f32(a)
    b, pb = _f32(b)
    return float(lib().fo_l2(pa, pb, a.size))


def dot(a, b) -> float:
    a, pa = _f32(a)
    b, pb = _f32(b)
    return float(lib().fo_dot(pa, pb, a.size))


def cosine(a, b) -> float:
    a, pa = _f32(a)
    b, pb = _f32(b)
    return float(lib().fo_cosine(pa, pb, a.size))


def l2_many(q, rows) -> np.ndarray:
    q, pq = _f32(q)
    rows, pr = _f32(rows)
    n, d = rows.shape
    out = np.empty(n, dtype=np.float32)
    lib().fo_l2_many(pq, pr, n, d, out.ctypes.data_as(_f32p))
    return out


# ---- top-k helpers (src/core/vector_ops.rs:12-32,180-260) ---------------------------------
def top_k_indices(scores, k, heap=False):
    s, ps = _f32(scores)
    out = np.empty(max(k, 1), dtype=np.uint32)
    f = lib().fo_top_k_indices_heap if heap else lib().fo_top_k_indices
    m = f(ps, s.size, k, out.ctypes.data_as(_u32p))
    return out[:m].tolist()


def streaming_top_k(scores, ids, k):
    s, ps = _f32(scores)
    i, pi = _u32(ids)
    oi = np.empty(max(k, 1), dtype=np.uint32)
    os_ = np.empty(max(k, 1), dtype=np.float32)
    m = lib().fo_streaming_top_k(ps, pi, s.size, k, oi.ctypes.data_as(_u32p),
                                 os_.ctypes.data_as(_f32p))
    return oi[:m].tolist(), os_[:m].tolist()


def merge_search_results(ids, dist, k):
    i, pi = _u32(ids)
    d, pd = _f32(dist)
    oi = np.empty(max(i.size, 1), dtype=np.uint32)
    od = np.empty(max(i.size, 1), dtype=np.float32)
    m = lib().fo_merge_search_results(pi, pd, i.size, k, oi.ctypes.data_as(_u32p),
                                      od.ctypes.data_as(_f32p))
    return oi[:m].tolist(), od[:m].tolist()


# ---- IVF / hybrid (src/ivf/core.rs, src/hybrid/core.rs) -----------------------------------
def find_nearest_centroid(x, centroids) -> int:
    x, px = _f32(x)
    c, pc = _f32(centroids)
    return int(lib().fo_find_nearest_centroid(px, pc, c.shape[0], c.shape[1]))


def assign(x, centroids) -> np.ndarray:
    x, px = _f32(x)
    c, pc = _f32(centroids)
    out = np.empty(x.shape[0], dtype=np.uint32)
    lib().fo_assign(px, x.shape[0], pc, c.shape[0], c.shape[1], out.ctypes.data_as(_u32p))
    return out


class IVF:
    """IVFIndex after set_trained(centroids) + insert(id, row) for every row
    (src/ivf/core.rs:431-455,509-520)."""

    def __init__(self, centroids, rows, ids=None, assign_=None):
        c, pc = _f32(centroids)
        x, px = _f32(rows)
        x = x.reshape(-1, c.shape[1])
        n = x.shape[0]
        if ids is None:
            ids = np.arange(n, dtype=np.uint32)
        i, pi = _u32(ids)
        self.dim = c.shape[1]
        self.nlist = c.shape[0]
        self.n = n
        self.centroids, self.rows, self.ids = c, x, i   # host copies (insertion order)
        if assign_ is None:
            self.assign = np.empty(n, dtype=np.uint32)
            self._h = lib().fo_ivf_build(pc, self.nlist, self.dim, px, pi, n,
                                         self.assign.ctypes.data_as(_u32p))
        else:
            self.assign, pa = _u32(assign_)
            self._h = lib().fo_ivf_build_assigned(pc, self.nlist, self.dim, px, pi, pa, n)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().fo_ivf_free(self._h)
                self._h = None
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass

    def list_len(self, l) -> int:
        return int(lib().fo_ivf_list_len(self._h, l))

    def coarse(self, q, nprobe):
        q, pq = _f32(q)
        ol = np.empty(max(nprobe, 1), dtype=np.uint32)
        od = np.empty(max(nprobe, 1), dtype=np.float32)
        m = lib().fo_ivf_coarse(self._h, pq, nprobe, ol.ctypes.data_as(_u32p),
                                od.ctypes.data_as(_f32p))
        return ol[:m].copy(), od[:m].copy()

    def search(self, q, k, nprobe, deleted=None, filter_bits=None, faithful=False):
        q, pq = _f32(q)
        oi = np.empty(max(k, 1), dtype=np.uint32)
        od = np.empty(max(k, 1), dtype=np.float32)
        _d, pd, nd = _bits(deleted)
        _f, pf, nf = _bits(filter_bits)
        if faithful:
            m = lib().fo_ivf_search_faithful(self._h, pq, k, nprobe, pd, nd,
                                             oi.ctypes.data_as(_u32p), od.ctypes.data_as(_f32p))
        else:
            m = lib().fo_ivf_search(self._h, pq, k, nprobe, pd, nd, pf, nf,
                                    oi.ctypes.data_as(_u32p), od.ctypes.data_as(_f32p))
        return oi[:m].copy(), od[:m].copy()


def flat_search(rows, ids, q, k, deleted=None, filter_bits=None):
    """Exact scan: ground truth / recent-tier replacement (src/hnsw/core.rs:398-467)."""
    rows, pr = _f32(rows)
    q, pq = _f32(q)
    n, d = rows.shape
    pi = None
    if ids is not None:
        ids, pi = _u32(ids)
    oi = np.empty(max(k, 1), dtype=np.uint32)
    od = np.empty(max(k, 1), dtype=np.float32)
    _d, pd, nd = _bits(deleted)
    _f, pf, nf = _bits(filter_bits)
    m = lib().fo_flat_search(pr, pi, n, d, pq, k, pd, nd, pf, nf, oi.ctypes.data_as(_u32p),
                             od.ctypes.data_as(_f32p))
    return oi[:m].copy(), od[:m].copy()


def hybrid_batch_search(ivf, flat_rows, flat_ids, q, k, nprobe, tiers=3, deleted=None,
                        filter_bits=None, threads=0, faithful=False):
    """Batched HybridIndex::search_with_config (src/hybrid/core.rs:425-486).
    Returns (ids [nq,k], dist [nq,k], count [nq])."""
    q, pq = _f32(q)
    nq, d = q.shape
    if flat_rows is None or len(flat_rows) == 0:
        fr, pfr, fi, pfi, fn = None, None, None, None, 0
    else:
        fr, pfr = _f32(flat_rows)
        fi, pfi = _u32(flat_ids)
        fn = fr.shape[0]
    oi = np.zeros((nq, max(k, 1)), dtype=np.uint32)
    od = np.full((nq, max(k, 1)), np.inf, dtype=np.float32)
    oc = np.zeros(nq, dtype=np.uint32)
    _d, pd, nd = _bits(deleted)
    _f, pf, nf = _bits(filter_bits)
    h = ivf._h if ivf is not None else None
    lib().fo_hybrid_batch_search(h, pfr, pfi, fn, d, pq, nq, k, nprobe, tiers, pd, nd, pf, nf,
                                 int(threads), int(bool(faithful)), oi.ctypes.data_as(_u32p),
                                 od.ctypes.data_as(_f32p), oc.ctypes.data_as(_u32p))
    return oi, od, oc


def hybrid_search_postfilter(ivf, flat_rows, flat_ids, q, k, nprobe, match_bits, tiers=3,
                             deleted=None):
    """HybridIndex::search_with_filter (src/hybrid/core.rs:513-549), one query."""
    q, pq = _f32(q)
    d = q.size
    if flat_rows is None or len(flat_rows) == 0:
        fr, pfr, fi, pfi, fn = None, None, None, None, 0
    else:
        fr, pfr = _f32(flat_rows)
        fi, pfi = _u32(flat_ids)
        fn = fr.shape[0]
    oi = np.empty(max(k, 1), dtype=np.uint32)
    od = np.empty(max(k, 1), dtype=np.float32)
    _d, pd, nd = _bits(deleted)
    _m, pm, nm = _bits(match_bits)
    h = ivf._h if ivf is not None else None
    m = lib().fo_hybrid_search_postfilter(h, pfr, pfi, fn, d, pq, k, nprobe, tiers, pd, nd, pm, nm,
                                          oi.ctypes.data_as(_u32p), od.ctypes.data_as(_f32p))
    return oi[:m].copy(), od[:m].copy()


COSINE, DOT = 1, 2


def flat_search_metric(rows, ids, queries, k, metric, deleted=None, filter_bits=None):
    """batch_cosine_similarity / dot scoring of every row + top_k_indices (src/core/vector_ops.rs:8-23) per
    query: (ids [nq,k], scores [nq,k], counts [nq]), best (largest) first, ties in input order."""
    x, px = _f32(rows)
    i, pi = _u32(ids)
    q, _ = _f32(queries)
    q = q.reshape(-1, x.shape[1])
    nq = q.shape[0]
    oi = np.full((nq, max(k, 1)), 0xFFFFFFFF, dtype=np.uint32)
    osc = np.full((nq, max(k, 1)), -np.inf, dtype=np.float32)
    oc = np.zeros(nq, dtype=np.uint32)
    _d, pd, nd = _bits(deleted)
    _f, pf, nf = _bits(filter_bits)
    for j in range(nq):
        qj = np.ascontiguousarray(q[j])
        oc[j] = lib().fo_flat_search_metric(px, pi, x.shape[0], x.shape[1], qj.ctypes.data_as(_f32p), k, metric, pd, nd,
                                            pf, nf, oi[j].ctypes.data_as(_u32p), osc[j].ctypes.data_as(_f32p))
    return oi, osc, oc


def recall(found, found_cnt, truth, truth_cnt, k) -> float:
    """evaluate_search_quality's recall (src/ivf/operations.rs:355-371)."""
    f, pf = _u32(found)
    fc, pfc = _u32(found_cnt)
    t, pt = _u32(truth)
    tc, ptc = _u32(truth_cnt)
    return float(lib().fo_recall(pf, pfc, pt, ptc, fc.size, k))


def num_threads() -> int:
    return int(lib().fo_num_threads())


# ---- k-means (src/ivf/core.rs:240-429) -----------------------------------------------------
def compute_error(data, centroids, assign_) -> float:
    x, px = _f32(data)
    c, pc = _f32(centroids)
    a, pa = _u32(assign_)
    return float(lib().fo_compute_error(px, x.shape[0], x.shape[1], pc, pa))


def update_centroids(data, assign_, centroids) -> np.ndarray:
    x, px = _f32(data)
    c = np.array(centroids, dtype=np.float32, order="C", copy=True)
    a, pa = _u32(assign_)
    lib().fo_update_centroids(px, x.shape[0], x.shape[1], c.shape[0], pa,
                              c.ctypes.data_as(_f32p))
    return c


def train_lloyd(data, init_centroids, max_iterations):
    """IVFIndex::train's Lloyd loop from shared initial centroids.
    Returns (centroids, assign, dict(iterations, converged, initial_error, final_error))."""
    x, px = _f32(data)
    c = np.array(init_centroids, dtype=np.float32, order="C", copy=True)
    n, d = x.shape
    a = np.empty(n, dtype=np.uint32)
    it = C.c_uint32()
    cv = C.c_uint32()
    e0 = C.c_float()
    e1 = C.c_float()
    lib().fo_train_lloyd(px, n, d, c.shape[0], max_iterations, c.ctypes.data_as(_f32p),
                         a.ctypes.data_as(_u32p), C.byref(it), C.byref(cv), C.byref(e0),
                         C.byref(e1))
    return c, a, dict(iterations=it.value, converged=bool(cv.value), initial_error=e0.value,
                      final_error=e1.value)


def retrain_lloyd(ivf, new_init_centroids, max_iterations):
    """IVFIndex::retrain (src/ivf/operations.rs:148-193) from shared initial centroids: collect the
    vectors of every inverted list (:157-162; the reference walks HashMaps in unspecified order —
    this restatement fixes the order to list-major, insertion order inside a list, the device's
    arena order), train, reinsert every vector into the list of its nearest new centroid
    (:186-188 -> insert, src/ivf/core.rs:431-455).
    Returns (new IVF over the same rows/ids, order [n] = indices of the old rows in training order,
    train dict)."""
    order = np.argsort(ivf.assign, kind="stable")
    x = ivf.rows[order]
    ids = ivf.ids[order]
    cent, a, res = train_lloyd(x, new_init_centroids, max_iterations)
    # reinsertion assigns against the FINAL centroids (a holds the last iteration's assignment,
    # taken before the last centroid update)
    return IVF(cent, x, ids), order, res


def kmeanspp_init(data, k, seed):
    x, px = _f32(data)
    n, d = x.shape
    c = np.zeros((k, d), dtype=np.float32)
    picked = np.zeros(k, dtype=np.uint32)
    m = lib().fo_kmeanspp_init(px, n, d, k, seed, c.ctypes.data_as(_f32p),
                               picked.ctypes.data_as(_u32p))
    return c[:m].copy(), picked[:m].copy()


def stdrng_stream(seed, n) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    lib().fo_stdrng_stream(seed, out.ctypes.data_as(_u32p), n)
    return out
