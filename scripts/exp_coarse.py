"""Coarse-step experiment: random unit-vector centroids / rows / queries, TC search, reports how many
queries fell back to the exact path (a proof failure or an `uncertain` selection) per nlist.

    python scripts/exp_coarse.py 1024 2048 4096
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fabstir_vectordb_b200 import Engine, _lib as L  # noqa: E402

D, NQ, NPROBE, K = 384, 1024, 32, 10
rng = np.random.default_rng(5)
for arg in sys.argv[1:]:
    nlist = int(arg)
    cents = rng.standard_normal((nlist, D)).astype(np.float32)
    cents /= np.linalg.norm(cents, axis=1, keepdims=True)
    rows = np.repeat(cents, 8, axis=0) + 0.3 * rng.standard_normal((nlist * 8, D)).astype(np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    q = rows[rng.choice(len(rows), NQ, replace=False)] + 0.05 * rng.standard_normal((NQ, D)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    eng = Engine(D, k_max=16)
    eng.set_option(L.OPT_SCAN_MODE, L.SCAN_TC)
    eng.set_centroids(cents)
    eng.ivf_add(rows, np.arange(len(rows), dtype=np.uint32))
    for rep in range(2):
        ids, dist, cnt = eng.search(q.astype(np.float32), K, NPROBE, tiers=L.TIER_HISTORICAL)
    st = eng.stats()
    print(f"nlist {nlist}: fallback queries {st.last_fallback_queries} of {NQ}, device_ms {st.last_device_ms:.3f}", flush=True)
    eng.close()
