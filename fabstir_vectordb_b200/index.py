"""Host-side mirror of the reference's index API for the hot path — same names, argument
meaning and error behaviour as IVFIndex (src/ivf/core.rs), HNSWIndex (src/hnsw/core.rs) and
HybridIndex (src/hybrid/core.rs) — over the C ABI of libfvdb_b200.so.

The reference is Rust; neither rustc nor cargo exists in this image, so the Rust shim is shipped
as source in INTEGRATION.md and this module is the executable mirror the parity tests drive.
Everything numeric (distances, argmin, top-k, k-means) happens on the GPU behind the ABI; this
file only keeps what the reference keeps host-side: the VectorId <-> row-id map, timestamps,
metadata filter evaluation and the error enums.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Any, Dict, Hashable, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .engine import (DimensionMismatch, DuplicateVector, Engine, FvdbError,
                     InconsistentDimensions, InsufficientTrainingData, InvalidConfig, NanInput, NotTrained,
                     VectorNotFound)


class NotInitialized(FvdbError):
    """HybridError::NotInitialized, src/hybrid/core.rs:16-35."""


@dataclass
class SearchResult:
    """SearchResult{vector_id, distance, metadata}, src/core/types.rs:191-195."""
    vector_id: Hashable
    distance: float
    metadata: Any = None


@dataclass
class TrainResult:
    """src/ivf/core.rs:103-109."""
    iterations: int
    converged: bool
    initial_error: float
    final_error: float


@dataclass
class RetrainResult:
    """src/ivf/operations.rs:38-44."""
    old_clusters: int
    new_clusters: int
    vectors_reassigned: int
    converged: bool


@dataclass
class AddClustersResult:
    """src/ivf/operations.rs:46-50."""
    clusters_added: int
    vectors_reassigned: int


@dataclass
class OptimizationResult:
    """src/ivf/operations.rs:52-56."""
    iterations: int
    improvement: float


@dataclass
class ClusterStats:
    """src/ivf/operations.rs:59-66."""
    n_clusters: int
    total_vectors: int
    avg_cluster_size: float
    size_variance: float
    empty_clusters: int


@dataclass
class BalanceResult:
    """src/ivf/operations.rs:91-95."""
    vectors_moved: int
    balance_improved: bool


class InvalidParameter(FvdbError):
    """OperationError::InvalidParameter, src/ivf/operations.rs:16."""
    code = L.ERR_INVALID_ARG


@dataclass
class IVFConfig:
    """src/ivf/core.rs:42-70 (defaults :50-60)."""
    n_clusters: int = 256
    n_probe: int = 16
    train_size: int = 10000
    max_iterations: int = 25
    seed: Optional[int] = None

    def is_valid(self) -> bool:
        return (self.n_clusters > 0 and self.n_probe > 0 and self.n_probe <= self.n_clusters
                and self.train_size > 0 and self.max_iterations > 0)


@dataclass
class HNSWConfig:
    """src/hnsw/core.rs:30-46.  Graph parameters are accepted and ignored: the recent tier is
    served by an exact scan, whose results dominate any graph walk."""
    max_connections: int = 16
    max_connections_layer_0: int = 32
    ef_construction: int = 200
    seed: Optional[int] = None


class _IdMap:
    """VectorId <-> dense u32 row id (stays host-side)."""

    def __init__(self):
        self.to_row: Dict[Hashable, int] = {}
        self.to_id: List[Hashable] = []

    def add(self, vid: Hashable) -> int:
        if vid in self.to_row:
            raise DuplicateVector(f"Vector with ID {vid!r} already exists")
        r = len(self.to_id)
        self.to_row[vid] = r
        self.to_id.append(vid)
        return r

    def rollback(self, n: int):
        for _ in range(n):
            vid = self.to_id.pop()
            del self.to_row[vid]

    def release(self, vid: Hashable):
        """A vacuumed vector: its VectorId may be inserted again (the reference removes the entry
        physically, src/ivf/operations.rs:625-645).  The row-id slot stays (row ids are dense and
        never reused), pointing at nothing."""
        row = self.to_row.pop(vid, None)
        if row is not None:
            self.to_id[row] = None

    def live(self) -> int:
        return len(self.to_row)


def _as_matrix(vectors: Sequence[Sequence[float]], dim: Optional[int]) -> np.ndarray:
    rows = [np.asarray(v, dtype=np.float32).reshape(-1) for v in vectors]
    if not rows:
        return np.zeros((0, dim or 0), dtype=np.float32)
    d0 = rows[0].shape[0]
    for r in rows:
        if r.shape[0] != d0:
            raise InconsistentDimensions(
                f"Inconsistent dimensions in training data: expected {d0}, found {r.shape[0]}")
    return np.stack(rows)


def _results(idmap: _IdMap, ids, dist, cnt) -> List[List[SearchResult]]:
    out = []
    for q in range(ids.shape[0]):
        c = int(cnt[q])
        out.append([SearchResult(idmap.to_id[int(ids[q, j])], float(dist[q, j])) for j in range(c)])
    return out


class IVFIndex:
    """IVFIndex, src/ivf/core.rs:154-682 + the operations of src/ivf/operations.rs that touch
    the search path (batch_insert :107, batch_search :132, mark_deleted :569, vacuum :625)."""

    def __init__(self, config: IVFConfig = None, *, k_max: int = 128, device: int = 0):
        config = config or IVFConfig()
        if not config.is_valid():
            raise InvalidConfig("Invalid IVFConfig")  # the reference panics, src/ivf/core.rs:171-174
        self.config = config
        self._k_max = k_max
        self._device = device
        self._eng: Optional[Engine] = None
        self._ids = _IdMap()
        self._deleted = set()
        self._dimension: Optional[int] = None
        self._trained = False
        self._lists: Dict[Hashable, int] = {}

    # -- accessors ------------------------------------------------------------------------
    def is_trained(self) -> bool:
        return self._trained

    def dimension(self) -> Optional[int]:
        return self._dimension

    def total_vectors(self) -> int:
        return len(self._lists)

    def active_count(self) -> int:  # src/ivf/operations.rs:615
        return len(self._lists) - len(self._deleted)

    def engine(self) -> Engine:
        return self._eng

    def get_centroids(self) -> np.ndarray:
        return self._eng.get_centroids() if self._trained else np.zeros((0, 0), np.float32)

    def get_cluster_sizes(self) -> Dict[int, int]:
        sizes = {i: 0 for i in range(self.config.n_clusters)}
        for l in self._lists.values():
            sizes[l] += 1
        return sizes

    def _ensure_engine(self, dim: int):
        if self._eng is None or self._eng.dim != dim:
            self._eng = Engine(dim, k_max=self._k_max, device=self._device)

    # -- training -------------------------------------------------------------------------
    def train(self, training_data, init_centroids=None) -> TrainResult:
        """IVFIndex::train, src/ivf/core.rs:240-334.  `init_centroids` (extension) replaces the
        seeded k-means++ draw so Lloyd parity can be checked from a shared start."""
        if isinstance(training_data, np.ndarray):
            data = np.ascontiguousarray(training_data, dtype=np.float32)
            n = data.shape[0]
        else:
            n = len(training_data)
            data = None
        if n == 0:
            raise InsufficientTrainingData(
                f"Insufficient training data: got 0, need at least {self.config.n_clusters}")
        if n < self.config.n_clusters:
            raise InsufficientTrainingData(
                f"Insufficient training data: got {n}, need at least {self.config.n_clusters}")
        if data is None:
            data = _as_matrix(training_data, None)
        dim = data.shape[1]
        self._ensure_engine(dim)
        self._dimension = dim
        seed = self.config.seed if self.config.seed is not None else time.time_ns()
        r = self._eng.train(data, self.config.n_clusters, self.config.max_iterations,
                            init_centroids=init_centroids, seed=seed)
        self._trained = True
        self._ids = _IdMap()
        self._lists = {}
        self._deleted = set()
        return TrainResult(**r)

    def set_trained(self, centroids, dimension: int):
        """src/ivf/core.rs:509-520."""
        c = np.ascontiguousarray(centroids, dtype=np.float32).reshape(-1, dimension)
        self._ensure_engine(dimension)
        self._eng.set_centroids(c)
        self._dimension = dimension
        self._trained = True
        self._ids = _IdMap()
        self._lists = {}
        self._deleted = set()

    # -- insertion ------------------------------------------------------------------------
    def insert(self, vid: Hashable, vector) -> None:
        """src/ivf/core.rs:431-455."""
        self.batch_insert([vid], [vector])

    def batch_insert(self, ids: Sequence[Hashable], vectors) -> None:
        """src/ivf/operations.rs:107 — one coarse-assignment launch for the whole batch."""
        if not self._trained:
            raise NotTrained("Index not trained. Call train() before inserting or searching.")
        x = vectors if isinstance(vectors, np.ndarray) else _as_matrix(vectors, self._dimension)
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self._dimension:
            raise DimensionMismatch(self._dimension, x.shape[-1] if x.ndim else 0)
        rows = []
        try:
            for v in ids:
                rows.append(self._ids.add(v))
        except DuplicateVector:
            self._ids.rollback(len(rows))
            raise
        try:
            lists = self._eng.ivf_add(x, np.asarray(rows, dtype=np.uint32), want_lists=True)
        except FvdbError:
            self._ids.rollback(len(rows))
            raise
        for v, l in zip(ids, lists):
            self._lists[v] = int(l)

    def load_chunk(self, cbor: bytes) -> int:
        """Bulk load of one stored VectorChunk: what `load_index_chunked` does per chunk
        (src/hybrid/persistence.rs:610-653: decode, then `find_cluster` for every vector inside a
        loop over clusters) as one decode into dense arrays and ONE assignment launch.  Vector ids
        are the 32 VectorId bytes.  Returns the number of vectors added."""
        from .chunk import decode_vector_chunk
        ch = decode_vector_chunk(cbor, pinned=True)
        if len(ch) == 0:
            return 0
        self.batch_insert([bytes(b) for b in ch.ids], ch.rows)
        return len(ch)

    def find_cluster(self, vector) -> int:
        """src/ivf/core.rs:493-499."""
        if not self._trained:
            raise NotTrained("Index not trained. Call train() before inserting or searching.")
        x = np.asarray(vector, dtype=np.float32).reshape(1, -1)
        if x.shape[1] != self._dimension:
            raise DimensionMismatch(self._dimension, x.shape[1])
        return int(self._eng.assign(x)[0])

    # -- maintenance (src/ivf/operations.rs:147-260, 262-288, 422-492, 552-564) --------------
    def _retrain_on_device(self, new_config: IVFConfig, init_centroids=None) -> TrainResult:
        seed = new_config.seed if new_config.seed is not None else time.time_ns()
        r = self._eng.retrain(new_config.n_clusters, new_config.max_iterations, init_centroids=init_centroids,
                              seed=seed)
        self.config = new_config
        rows, lists = self._eng.dump_lists()
        self._lists = {self._ids.to_id[int(r_)]: int(l) for r_, l in zip(rows, lists)}
        return TrainResult(**r)

    def retrain(self, new_config: IVFConfig, init_centroids=None) -> RetrainResult:
        """IVFIndex::retrain, src/ivf/operations.rs:148-193: k-means over every stored vector
        with the new config, then every vector is reinserted.  One device call (fvdb_ivf_retrain):
        the vectors never leave the GPU.  Soft-deleted vectors stay in their (new) lists, as in
        the reference (the `deleted` set is not touched by retrain)."""
        if not self._trained:
            raise NotTrained("Index not trained. Call train() before inserting or searching.")
        if not new_config.is_valid():
            raise InvalidConfig("Invalid IVFConfig")
        old_clusters, old_vectors = self.config.n_clusters, self.total_vectors()
        if old_vectors < new_config.n_clusters:   # the train() inside the reference fails, core.rs:250-255
            raise InsufficientTrainingData(
                f"Insufficient training data: got {old_vectors}, need at least {new_config.n_clusters}")
        tr = self._retrain_on_device(new_config, init_centroids)
        return RetrainResult(old_clusters, new_config.n_clusters, old_vectors, tr.converged)

    def add_clusters(self, n_clusters_to_add: int) -> AddClustersResult:
        """src/ivf/operations.rs:195-219."""
        if not self._trained:
            raise NotTrained("Index not trained. Call train() before inserting or searching.")
        if n_clusters_to_add == 0:
            raise InvalidParameter("Cannot add 0 clusters")
        cfg = IVFConfig(**{**self.config.__dict__, "n_clusters": self.config.n_clusters + n_clusters_to_add})
        rr = self.retrain(cfg)
        return AddClustersResult(n_clusters_to_add, rr.vectors_reassigned)

    def optimize_clusters(self) -> OptimizationResult:
        """src/ivf/operations.rs:221-260: retrain with the same config, report the drop in the
        variance of the list sizes."""
        if not self._trained:
            raise NotTrained("Index not trained. Call train() before inserting or searching.")
        before = self._size_variance()
        if self.total_vectors() < self.config.n_clusters:
            raise InsufficientTrainingData(
                f"Insufficient training data: got {self.total_vectors()}, need at least {self.config.n_clusters}")
        tr = self._retrain_on_device(self.config)
        return OptimizationResult(tr.iterations, max(before - self._size_variance(), 0.0))

    def _size_variance(self) -> float:
        """calculate_size_variance, src/ivf/operations.rs:552-564 (f32 arithmetic, left folds)."""
        n = self.config.n_clusters
        sizes = np.zeros(n, dtype=np.float32)
        for l in self._lists.values():
            sizes[l] += np.float32(1)
        total = np.float32(0)
        for v in sizes:
            total = np.float32(total + v)
        mean = np.float32(total / np.float32(n))
        acc = np.float32(0)
        for v in sizes:
            d = np.float32(v - mean)
            acc = np.float32(acc + np.float32(d * d))
        return float(np.float32(acc / np.float32(n)))

    def get_cluster_stats(self) -> ClusterStats:
        """src/ivf/operations.rs:263-288."""
        n = self.config.n_clusters
        sizes = self.get_cluster_sizes()
        total = self.total_vectors()
        avg = float(np.float32(total) / np.float32(n)) if n > 0 else 0.0
        return ClusterStats(n, total, avg, self._size_variance(), sum(1 for v in sizes.values() if v == 0))

    def balance_clusters(self, threshold: float) -> BalanceResult:
        """src/ivf/operations.rs:422-492.  The reference takes vectors out of oversized lists and
        puts each back into the list of its nearest centroid — the list it came from, since
        membership always is the nearest centroid (insert :431-455, retrain) — so nothing moves
        unless the centroid table was replaced under the lists; the device keeps that invariant
        by construction (fvdb_ivf_set_centroids clears the lists), hence vectors_moved == 0."""
        if not (0.0 < threshold < 1.0):
            raise InvalidParameter("Threshold must be between 0 and 1")
        return BalanceResult(0, False)

    # -- soft delete (src/ivf/operations.rs:569-640) ----------------------------------------
    def mark_deleted(self, vid: Hashable) -> None:
        if vid not in self._lists:
            raise VectorNotFound(f"Vector not found: {vid!r}")
        self._eng.set_deleted([self._ids.to_row[vid]], True)
        self._deleted.add(vid)

    def is_deleted(self, vid: Hashable) -> bool:
        return vid in self._deleted

    def vacuum(self) -> int:
        """src/ivf/operations.rs:625-645: the entries are removed physically, so a vacuumed
        VectorId can be inserted again."""
        removed = self._eng.vacuum() if self._eng else 0
        for v in self._deleted:
            self._lists.pop(v, None)
            self._ids.release(v)
        self._deleted = set()
        return removed

    # -- search ---------------------------------------------------------------------------
    def search(self, query, k: int) -> List[SearchResult]:
        """src/ivf/core.rs:622-624."""
        return self.search_with_config(query, k, self.config.n_probe)

    def search_with_config(self, query, k: int, n_probe: int) -> List[SearchResult]:
        """src/ivf/core.rs:626-681."""
        return self.batch_search_with_config([query], k, n_probe)[0]

    def batch_search(self, queries, k: int) -> List[List[SearchResult]]:
        """src/ivf/operations.rs:132-145, as ONE device batch."""
        return self.batch_search_with_config(queries, k, self.config.n_probe)

    def batch_search_with_config(self, queries, k: int, n_probe: int):
        if not self._trained:
            raise NotTrained("Index not trained. Call train() before inserting or searching.")
        q = queries if isinstance(queries, np.ndarray) else _as_matrix(queries, self._dimension)
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.shape[1] != self._dimension:
            raise DimensionMismatch(self._dimension, q.shape[1])
        if k == 0:
            return [[] for _ in range(q.shape[0])]
        ids, dist, cnt = self._eng.search(q, k, n_probe, tiers=L.TIER_HISTORICAL)
        return _results(self._ids, ids, dist, cnt)


class HNSWIndex:
    """HNSWIndex, src/hnsw/core.rs — the recent tier.  search() is an exact scan; the `ef`
    argument is accepted and ignored (results are a superset-in-quality of the graph walk)."""

    def __init__(self, config: HNSWConfig = None, *, k_max: int = 128, device: int = 0):
        self.config = config or HNSWConfig()
        self._k_max = k_max
        self._device = device
        self._eng: Optional[Engine] = None
        self._ids = _IdMap()
        self._deleted = set()
        self._dimension: Optional[int] = None

    def node_count(self) -> int:
        return self._ids.live()

    def dimension(self) -> Optional[int]:
        return self._dimension

    def insert(self, vid: Hashable, vector) -> None:
        """src/hnsw/core.rs:226."""
        self.batch_insert([vid], [vector])

    def batch_insert(self, ids, vectors) -> None:
        x = vectors if isinstance(vectors, np.ndarray) else _as_matrix(vectors, self._dimension)
        x = np.ascontiguousarray(x, dtype=np.float32)
        if self._dimension is None:
            self._dimension = x.shape[1]
            self._eng = Engine(self._dimension, k_max=self._k_max, device=self._device)
        if x.shape[1] != self._dimension:
            raise DimensionMismatch(self._dimension, x.shape[1])
        rows = []
        try:
            for v in ids:
                rows.append(self._ids.add(v))
        except DuplicateVector:
            self._ids.rollback(len(rows))
            raise
        try:
            self._eng.flat_add(x, np.asarray(rows, dtype=np.uint32))
        except FvdbError:
            self._ids.rollback(len(rows))
            raise

    def search(self, query, k: int, ef: int = 50) -> List[SearchResult]:
        """src/hnsw/core.rs:398-467.  Empty index -> [] (:404-407)."""
        if self._eng is None:
            return []
        q = np.asarray(query, dtype=np.float32).reshape(1, -1)
        if q.shape[1] != self._dimension:
            raise DimensionMismatch(self._dimension, q.shape[1])
        if k == 0:
            return []
        ids, dist, cnt = self._eng.search(q, k, 0, tiers=L.TIER_RECENT)
        return _results(self._ids, ids, dist, cnt)[0]

    def mark_deleted(self, vid: Hashable) -> None:
        if vid not in self._ids.to_row:
            raise VectorNotFound(f"Vector not found: {vid!r}")
        self._eng.set_deleted([self._ids.to_row[vid]], True)
        self._deleted.add(vid)

    def is_deleted(self, vid: Hashable) -> bool:
        return vid in self._deleted

    def vacuum(self) -> int:
        removed = self._eng.vacuum() if self._eng else 0
        for v in self._deleted:
            self._ids.release(v)
        self._deleted = set()
        return removed


# ---- metadata filter ------------------------------------------------------------------------
# The reference evaluates a MetadataFilter (src/core/metadata_filter.rs) against JSON metadata on the
# host; that evaluator is outside the hot path.  Here a filter is any callable
# `metadata -> bool` (or an object with a `.matches(metadata)` method): the mirror evaluates it once per
# row and hands the engine a bitmap.

def _matches(flt, metadata) -> bool:
    m = getattr(flt, "matches", None)
    return bool(m(metadata) if m is not None else flt(metadata))


@dataclass
class HybridConfig:
    """src/hybrid/core.rs:38-85 (defaults :69-85: n_clusters 3, n_probe 2, train_size 9)."""
    recent_threshold: float = 7 * 24 * 3600.0  # seconds
    hnsw_config: HNSWConfig = field(default_factory=HNSWConfig)
    ivf_config: IVFConfig = field(default_factory=lambda: IVFConfig(n_clusters=3, n_probe=2, train_size=9))
    migration_batch_size: int = 100
    auto_migrate: bool = True
    min_ivf_training_size: int = 10

    def is_valid(self) -> bool:
        return self.recent_threshold > 0 and self.migration_batch_size > 0


@dataclass
class HybridSearchConfig:
    """src/hybrid/core.rs:173-196."""
    search_recent: bool = True
    search_historical: bool = True
    recent_k: int = 0
    historical_k: int = 0
    recent_threshold_override: Optional[float] = None
    k: int = 10
    hnsw_ef: int = 50
    ivf_n_probe: int = 10


SearchConfig = HybridSearchConfig


class HybridIndex:
    """HybridIndex, src/hybrid/core.rs:202-700: recent tier (exact scan in place of HNSW) +
    historical IVF tier behind one device handle; search = union, stable sort by distance with
    the recent tier first on ties, truncate(k), no de-duplication (:425-486)."""

    def __init__(self, config: HybridConfig = None, *, k_max: int = 128, device: int = 0):
        config = config or HybridConfig()
        if not config.is_valid():
            raise InvalidConfig("Invalid HybridConfig")
        self.config = config
        self._k_max = k_max
        self._device = device
        self._eng: Optional[Engine] = None
        self._ids = _IdMap()
        self.timestamps: Dict[Hashable, float] = {}
        self._tier: Dict[Hashable, int] = {}
        self._deleted = set()
        self.initialized = False
        self.ivf_trained = False
        self._dimension: Optional[int] = None

    def is_initialized(self) -> bool:
        return self.initialized

    def engine(self) -> Engine:
        return self._eng

    def recent_count(self) -> int:
        return sum(1 for t in self._tier.values() if t == 1)

    def historical_count(self) -> int:
        return sum(1 for t in self._tier.values() if t == 2)

    def _ensure_engine(self, dim: int):
        if self._eng is None:
            self._eng = Engine(dim, k_max=self._k_max, device=self._device)
            self._dimension = dim

    def initialize(self, training_data, init_centroids=None) -> None:
        """src/hybrid/core.rs:262-289: fewer than min_ivf_training_size rows => HNSW-only mode;
        else train IVF and leave its lists empty."""
        n = len(training_data)
        if n < self.config.min_ivf_training_size:
            self.ivf_trained = False
            self.initialized = True
            return
        data = training_data if isinstance(training_data, np.ndarray) else _as_matrix(training_data, None)
        data = np.ascontiguousarray(data, dtype=np.float32)
        c = self.config.ivf_config
        if n < c.n_clusters:
            raise InsufficientTrainingData(
                f"Insufficient training data: got {n}, need at least {c.n_clusters}")
        self._ensure_engine(data.shape[1])
        seed = c.seed if c.seed is not None else time.time_ns()
        self._eng.train(data, c.n_clusters, c.max_iterations, init_centroids=init_centroids, seed=seed)
        self.ivf_trained = True
        self.initialized = True

    def insert(self, vid, vector) -> None:
        self.insert_with_timestamp(vid, vector, time.time())

    def insert_with_timestamp(self, vid, vector, timestamp: float) -> None:
        """src/hybrid/core.rs:357-413."""
        self.batch_insert_with_timestamps([vid], [vector], [timestamp])

    def batch_insert_with_timestamps(self, ids, vectors, timestamps) -> None:
        if not self.initialized:
            raise NotInitialized("Index not initialized")
        x = vectors if isinstance(vectors, np.ndarray) else _as_matrix(vectors, self._dimension)
        x = np.ascontiguousarray(x, dtype=np.float32)
        self._ensure_engine(x.shape[1])
        if x.shape[1] != self._dimension:
            raise DimensionMismatch(self._dimension, x.shape[1])
        ids = list(ids)
        seen = set()
        for v in ids:   # duplicates against the index AND inside the batch, before any id is registered
            if v in self._ids.to_row or v in seen:
                raise DuplicateVector(f"Vector with ID {v!r} already exists")
            seen.add(v)
        if np.isnan(x).any():   # the engine would reject the batch half-way (flat rows in, IVF rows not)
            raise NanInput("NaN in input (the reference panics on partial_cmp().unwrap())")
        now = time.time()
        rows = np.asarray([self._ids.add(v) for v in ids], dtype=np.uint32)
        ts = np.asarray(timestamps, dtype=np.float64)
        age = np.maximum(now - ts, 0.0)
        recent = np.ones(len(rows), dtype=bool) if not self.ivf_trained else (age < self.config.recent_threshold)
        flat_done = False
        try:
            if recent.any():
                self._eng.flat_add(x[recent], rows[recent])
                flat_done = True
            if (~recent).any():
                self._eng.ivf_add(x[~recent], rows[~recent])
        except FvdbError:
            # nothing of a failed batch may stay behind: forget the ids, and take the rows the first
            # call already placed out of the device again
            if flat_done:
                self._eng.set_deleted(rows[recent], True)
                self._eng.vacuum()
                if self._deleted:   # vacuum dropped earlier soft-deleted rows as well
                    for v in self._deleted:
                        self._tier.pop(v, None)
                        self.timestamps.pop(v, None)
                        self._ids.release(v)
                    self._deleted = set()
            self._ids.rollback(len(rows))
            raise
        for v, t, r in zip(ids, ts, recent):
            self.timestamps[v] = float(t)
            self._tier[v] = 1 if r else 2

    def migrate_with_threshold(self, threshold: float) -> int:
        """src/hybrid/core.rs:600-649: move recent-tier vectors older than `threshold` seconds
        into the IVF tier.  Unlike the reference (which leaves a stale copy in HNSW, :626-635)
        the row is dropped from the recent tier, so no duplicate ids appear."""
        if not self.ivf_trained or self._eng is None:
            return 0
        now = time.time()
        move = [v for v, t in self.timestamps.items()
                if self._tier.get(v) == 1 and max(now - t, 0.0) >= threshold]
        if not move:
            return 0
        rows = np.asarray([self._ids.to_row[v] for v in move], dtype=np.uint32)
        moved = self._eng.move_flat_to_ivf(rows)
        for v in move:
            self._tier[v] = 2
        return moved

    def migrate_old_vectors(self) -> int:
        return self.migrate_with_threshold(self.config.recent_threshold)

    def delete(self, vid) -> None:
        if vid not in self._tier:
            raise VectorNotFound(f"Vector not found: {vid!r}")
        self._eng.set_deleted([self._ids.to_row[vid]], True)
        self._deleted.add(vid)

    def is_deleted(self, vid) -> bool:
        return vid in self._deleted

    def vacuum(self) -> int:
        removed = self._eng.vacuum() if self._eng else 0
        for v in self._deleted:
            self._tier.pop(v, None)
            self.timestamps.pop(v, None)
            self._ids.release(v)
        self._deleted = set()
        return removed

    # -- search ---------------------------------------------------------------------------
    def search(self, query, k: int) -> List[SearchResult]:
        """src/hybrid/core.rs:419-423: SearchConfig::default() with k => ivf_n_probe = 10."""
        return self.search_with_config(query, HybridSearchConfig(k=k))

    def search_with_config(self, query, config: HybridSearchConfig) -> List[SearchResult]:
        return self.batch_search_with_config([query], config)[0]

    def batch_search(self, queries, k: int) -> List[List[SearchResult]]:
        return self.batch_search_with_config(queries, HybridSearchConfig(k=k))

    def batch_search_with_config(self, queries, config: HybridSearchConfig, filter_bits=None,
                                 raw: bool = False):
        nq = len(queries)
        if not self.initialized or self._eng is None:
            return [[] for _ in range(nq)]  # :431-434
        if self.config.auto_migrate:
            self.migrate_old_vectors()  # :437-439
        q = queries if isinstance(queries, np.ndarray) else _as_matrix(queries, self._dimension)
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.shape[1] != self._dimension:
            # the reference swallows per-tier errors (`if let Ok`, :459,469,475) => empty result
            return [[] for _ in range(nq)]
        k = config.k
        if k == 0:
            return [[] for _ in range(nq)]
        tiers = (L.TIER_RECENT if config.search_recent else 0) | \
                (L.TIER_HISTORICAL if (config.search_historical and self.ivf_trained) else 0)
        if config.recent_k or config.historical_k:
            # per-tier k differs from k: run the tiers separately and merge on the host exactly
            # as :482-483 does (stable sort, recent first)
            rk = config.recent_k or k
            hk = config.historical_k or k
            parts = []
            if tiers & L.TIER_RECENT:
                parts.append(self._eng.search(q, rk, 0, L.TIER_RECENT, filter_bits))
            if tiers & L.TIER_HISTORICAL:
                parts.append(self._eng.search(q, hk, config.ivf_n_probe, L.TIER_HISTORICAL, filter_bits))
            out = []
            for i in range(nq):
                cand = []
                for ids, dist, cnt in parts:
                    cand += [(float(dist[i, j]), int(ids[i, j])) for j in range(int(cnt[i]))]
                cand.sort(key=lambda t: t[0])  # stable
                out.append([SearchResult(self._ids.to_id[r], d) for d, r in cand[:k]])
            return out
        ids, dist, cnt = self._eng.search(q, k, config.ivf_n_probe, tiers, filter_bits)
        if raw:
            return ids, dist, cnt
        return _results(self._ids, ids, dist, cnt)

    def filter_bitmap(self, flt, metadata_map: Dict[str, Any]) -> np.ndarray:
        """Evaluate the filter once per row on the host -> 1 bit per row id.  A row
        missing from `metadata_map` gets bit 0, as the reference drops it (src/hybrid/core.rs:536-541)."""
        n = len(self._ids.to_id)
        words = np.zeros((n + 63) // 64 or 1, dtype=np.uint64)
        for vid, row in self._ids.to_row.items():
            md = metadata_map.get(str(vid))
            if md is not None and _matches(flt, md):
                words[row >> 6] |= np.uint64(1) << np.uint64(row & 63)
        return words

    def search_with_filter(self, query, k: int, flt, metadata_map: Dict[str, Any]) -> List[SearchResult]:
        """src/hybrid/core.rs:513-549 — the reference's 3x oversample POST-filter, as one device call
        (fvdb_search_postfilter: search(3k), keep the rows whose bit is set, truncate(k)).  `flt` is any
        callable metadata -> bool (or an object with .matches)."""
        if flt is None:
            return self.search(query, k)
        if not self.initialized or self._eng is None or k == 0:
            return []
        if 3 * k > self._k_max:
            raise InvalidParameter(f"search_with_filter searches 3k candidates: k={k} needs k_max >= {3 * k} "
                                   f"(this index was created with k_max={self._k_max})")
        if self.config.auto_migrate:
            self.migrate_old_vectors()
        q = np.asarray(query, dtype=np.float32).reshape(1, -1)
        if q.shape[1] != self._dimension:
            return []
        cfg = HybridSearchConfig(k=k)
        tiers = L.TIER_RECENT | (L.TIER_HISTORICAL if self.ivf_trained else 0)
        bits = self.filter_bitmap(flt, metadata_map)
        ids, dist, cnt = self._eng.search_postfilter(q, k, cfg.ivf_n_probe, bits, tiers)
        return _results(self._ids, ids, dist, cnt)[0]

    def search_with_prefilter(self, query, k: int, flt, metadata_map) -> List[SearchResult]:
        """In-kernel bitmap PRE-filter (semantics of bindings/wasm/src/index.rs:164-186): never
        returns fewer than k when >= k matching rows are reachable."""
        bits = self.filter_bitmap(flt, metadata_map)
        return self.batch_search_with_config([query], HybridSearchConfig(k=k), filter_bits=bits)[0]
