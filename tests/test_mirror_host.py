"""CPU tests of the host mirror's bookkeeping (fabstir_vectordb_b200/index.py: VectorId <-> row id map, timestamps,
tiers, the deleted set, roll-backs) with the engine replaced by an oracle-backed stand-in.  The numeric parity of the
real engine is the GPU suite's job (tests/test_gpu_mirror.py runs the same scenarios against libfvdb_b200)."""
import time

import numpy as np
import pytest

import oracle as O
import fabstir_vectordb_b200.index as IDX
from fabstir_vectordb_b200 import (DuplicateVector, HNSWIndex, HybridConfig, HybridIndex, IVFConfig, IVFIndex,
                                   InvalidParameter, NanInput, VectorNotFound, _lib as L, synth)

D = 16


class OracleEngine:
    """The subset of Engine the mirrors use, served by the CPU oracle."""

    def __init__(self, dim, k_max=128, device=0, metric=0):
        self.dim, self.k_max = dim, k_max
        self.cents = None
        self.ivf_x, self.ivf_ids = np.zeros((0, dim), np.float32), np.zeros(0, np.uint32)
        self.flat_x, self.flat_ids = np.zeros((0, dim), np.float32), np.zeros(0, np.uint32)
        self.dead = set()

    @staticmethod
    def _check(x):
        if np.isnan(x).any():
            raise NanInput("NaN in input")

    def train(self, data, nlist, max_iterations, init_centroids=None, seed=0):
        data = np.ascontiguousarray(data, np.float32)
        init = init_centroids if init_centroids is not None else O.kmeanspp_init(data, nlist, seed)[0]
        self.cents, _, res = O.train_lloyd(data, init, max_iterations)
        self.ivf_x, self.ivf_ids = np.zeros((0, self.dim), np.float32), np.zeros(0, np.uint32)
        return res

    def ivf_add(self, x, row_ids, want_lists=False):
        x = np.ascontiguousarray(x, np.float32)
        self._check(x)
        self.ivf_x = np.concatenate([self.ivf_x, x])
        self.ivf_ids = np.concatenate([self.ivf_ids, np.asarray(row_ids, np.uint32)])
        return O.assign(x, self.cents) if want_lists else None

    def flat_add(self, x, row_ids):
        x = np.ascontiguousarray(x, np.float32)
        self._check(x)
        self.flat_x = np.concatenate([self.flat_x, x])
        self.flat_ids = np.concatenate([self.flat_ids, np.asarray(row_ids, np.uint32)])

    def set_deleted(self, row_ids, deleted=True):
        for r in np.asarray(row_ids).tolist():
            (self.dead.add if deleted else self.dead.discard)(int(r))

    def vacuum(self):
        n0 = len(self.ivf_ids) + len(self.flat_ids)
        ki = ~np.isin(self.ivf_ids, list(self.dead))
        kf = ~np.isin(self.flat_ids, list(self.dead))
        self.ivf_x, self.ivf_ids, self.flat_x, self.flat_ids = self.ivf_x[ki], self.ivf_ids[ki], self.flat_x[kf], self.flat_ids[kf]
        self.dead = set()
        return n0 - len(self.ivf_ids) - len(self.flat_ids)

    def _bitmap(self):
        n = int(max([0] + self.ivf_ids.tolist() + self.flat_ids.tolist())) + 1
        return O.make_bitmap(n, sorted(self.dead)) if self.dead else None

    def _ivf(self):
        return O.IVF(self.cents, self.ivf_x, self.ivf_ids) if self.cents is not None and len(self.ivf_ids) else None

    def search(self, q, k, nprobe, tiers=L.TIER_BOTH, filter_bits=None, out=None):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, self.dim)
        ivf = self._ivf() if tiers & L.TIER_HISTORICAL else None
        t = (1 if (tiers & L.TIER_RECENT) and len(self.flat_ids) else 0) | (2 if ivf is not None else 0)
        if t == 0:
            return (np.zeros((len(q), k), np.uint32), np.zeros((len(q), k), np.float32), np.zeros(len(q), np.uint32))
        return O.hybrid_batch_search(ivf, self.flat_x, self.flat_ids, q, k, nprobe, tiers=t, deleted=self._bitmap(),
                                     filter_bits=filter_bits)

    def search_postfilter(self, q, k, nprobe, keep_bits, tiers=L.TIER_BOTH):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, self.dim)
        ids = np.full((len(q), k), 0xFFFFFFFF, np.uint32)
        dist = np.full((len(q), k), np.inf, np.float32)
        cnt = np.zeros(len(q), np.uint32)
        for i in range(len(q)):
            a, b = O.hybrid_search_postfilter(self._ivf(), self.flat_x, self.flat_ids, q[i], k, nprobe, keep_bits,
                                              tiers=3 if self._ivf() is not None else 1, deleted=self._bitmap())
            ids[i, :len(a)], dist[i, :len(a)], cnt[i] = a, b, len(a)
        return ids, dist, cnt

    def move_flat_to_ivf(self, row_ids):
        m = np.isin(self.flat_ids, np.asarray(row_ids, np.uint32))
        self.ivf_x = np.concatenate([self.ivf_x, self.flat_x[m]])
        self.ivf_ids = np.concatenate([self.ivf_ids, self.flat_ids[m]])
        self.flat_x, self.flat_ids = self.flat_x[~m], self.flat_ids[~m]
        return int(m.sum())


@pytest.fixture(autouse=True)
def _oracle_engine(monkeypatch):
    monkeypatch.setattr(IDX, "Engine", OracleEngine)


def _rows(n, seed):
    return synth.rows(0, n, D, 8, 0.7, seed)


def test_vacuum_releases_ids_on_all_three_mirrors():
    x = _rows(300, 5)
    ivf = IVFIndex(IVFConfig(n_clusters=4, n_probe=4, train_size=100, max_iterations=3), k_max=16)
    ivf.train(x[:100], init_centroids=x[:4].copy())
    ivf.batch_insert(list(range(50)), x[:50])
    ivf.mark_deleted(7)
    assert ivf.is_deleted(7) and ivf.active_count() == 49
    assert ivf.vacuum() == 1 and ivf.total_vectors() == 49
    with pytest.raises(VectorNotFound):
        ivf.mark_deleted(7)
    ivf.insert(7, x[200])                                   # the reference removed the entry physically: re-insert works
    assert ivf.search_with_config(x[200], 1, 4)[0].vector_id == 7
    hn = HNSWIndex(k_max=16)
    hn.batch_insert(["a", "b", "c"], x[:3])
    hn.mark_deleted("b")
    assert hn.vacuum() == 1 and hn.node_count() == 2
    hn.insert("b", x[201])
    assert hn.node_count() == 3 and hn.search(x[201], 1)[0].vector_id == "b"
    hy = HybridIndex(HybridConfig(ivf_config=IVFConfig(n_clusters=4, n_probe=4, train_size=100, max_iterations=3),
                                  auto_migrate=False), k_max=16)
    hy.initialize(x[:100], init_centroids=x[:4].copy())
    now = time.time()
    hy.batch_insert_with_timestamps(list(range(20)), x[:20], [now - 1e7] * 10 + [now] * 10)
    hy.delete(3)
    hy.delete(15)
    assert hy.vacuum() == 2 and 3 not in hy.timestamps and 15 not in hy.timestamps
    hy.insert_with_timestamp(3, x[202], now)
    assert hy.recent_count() == 10 and hy.historical_count() == 9
    assert hy.search(x[202], 1)[0].vector_id == 3


def test_failed_batch_insert_rolls_everything_back():
    x = _rows(100, 9)
    hy = HybridIndex(HybridConfig(ivf_config=IVFConfig(n_clusters=4, n_probe=4, train_size=50, max_iterations=3),
                                  auto_migrate=False), k_max=16)
    hy.initialize(x[:50], init_centroids=x[:4].copy())
    now = time.time()
    hy.batch_insert_with_timestamps(["a"], x[:1], [now])
    with pytest.raises(DuplicateVector):
        hy.batch_insert_with_timestamps(["b", "c", "b"], x[1:4], [now] * 3)      # duplicate inside the batch
    with pytest.raises(DuplicateVector):
        hy.batch_insert_with_timestamps(["d", "a"], x[4:6], [now] * 2)           # duplicate against the index
    bad = x[6:8].copy()
    bad[1, 2] = np.nan
    with pytest.raises(NanInput):
        hy.batch_insert_with_timestamps(["e", "f"], bad, [now, now - 1e7])
    hy.batch_insert_with_timestamps(["b", "c", "d", "e", "f"], x[10:15], [now] * 5)   # every id is free again
    assert hy.recent_count() == 6
    hn = HNSWIndex(k_max=16)
    hn.batch_insert(["p"], x[:1])
    with pytest.raises(NanInput):
        hn.batch_insert(["q", "r"], bad)
    hn.batch_insert(["q", "r"], x[20:22])
    assert hn.node_count() == 3


def test_search_with_filter_maps_metadata_to_the_keep_bitmap():
    x = _rows(400, 13)
    hy = HybridIndex(HybridConfig(ivf_config=IVFConfig(n_clusters=4, n_probe=4, train_size=100, max_iterations=3),
                                  auto_migrate=False), k_max=16)
    hy.initialize(x[:100], init_centroids=x[:4].copy())
    now = time.time()
    vids = [f"v{i}" for i in range(400)]
    hy.batch_insert_with_timestamps(vids, x, [now - 1e7] * 300 + [now] * 100)
    meta = {v: {"even": i % 2 == 0} for i, v in enumerate(vids) if i % 7}         # every 7th row has no metadata
    flt = lambda md: md["even"]                                                    # noqa: E731
    bits = hy.filter_bitmap(flt, meta)
    want_rows = [i for i in range(400) if i % 7 and i % 2 == 0]
    assert np.array_equal(bits[:7], O.make_bitmap(448, want_rows)[:7])
    got = hy.search_with_filter(x[10] + np.float32(0.01), 5, flt, meta)
    assert all(int(g.vector_id[1:]) % 2 == 0 and int(g.vector_id[1:]) % 7 for g in got) and len(got) <= 5
    assert [g.distance for g in got] == sorted(g.distance for g in got)
    assert [r.vector_id for r in hy.search_with_filter(x[10], 5, None, meta)] == [r.vector_id for r in hy.search(x[10], 5)]
    with pytest.raises(InvalidParameter):
        hy.search_with_filter(x[10], 6, flt, meta)                                 # 3k = 18 > k_max = 16
