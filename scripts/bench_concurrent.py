"""Single-query search calls from T concurrent OS threads (BASELINE.json configs[0] shape: 100K x 384,
30 % recent tier / 70 % IVF, nlist 256, nprobe 16, k = 10 — the way the reference's callers use
HybridIndex::search, one query per call): calls/s through fvdb_search with the submission queue
(concurrent calls coalesced into one device batch) and with FVDB_OPT_COALESCE = 0 (serialised).
The Python threads release the GIL inside the ctypes call; the per-call Python overhead
(~10 us of argument marshalling under the GIL) is part of the number.

    python scripts/bench_concurrent.py [threads ...]
"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fabstir_vectordb_b200 import Engine, _lib as L, synth  # noqa: E402

N, D, NLIST, NPROBE, K = 100_000, 384, 256, 16, 10
N_RECENT = 30_000
x = synth.rows(0, N, D, 4 * NLIST, 1.0, 1234)
q = synth.queries(0, 4096, D, N, 4 * NLIST, 1.0, 1234, synth.default_qnoise(D, 1.0), 5678)
eng = Engine(D, k_max=16)
eng.set_option(L.OPT_SCAN_MODE, L.SCAN_TC)
res = eng.train(x[:: N // (NLIST * 40)][: NLIST * 40], NLIST, 10, seed=7)
eng.ivf_add(x[N_RECENT:], np.arange(N_RECENT, N, dtype=np.uint32))
eng.flat_add(x[:N_RECENT], np.arange(N_RECENT, dtype=np.uint32))
eng.search(q[:8], K, NPROBE, tiers=L.TIER_BOTH)
ref = eng.search(q[:256], K, NPROBE, tiers=L.TIER_BOTH)
SECONDS = float(os.environ.get("FVDB_CONC_SECONDS", 1.5))


def run(threads, coalesce):
    eng.set_option(L.OPT_COALESCE, 1 if coalesce else 0)
    counts = [0] * threads
    batch_calls = []
    stop = time.perf_counter() + SECONDS
    ok = [True]

    def worker(t):
        out = (np.empty((1, K), np.uint32), np.empty((1, K), np.float32), np.zeros(1, np.uint32))
        i = t
        while time.perf_counter() < stop:
            qi = i % 256
            ids, dist, cnt = eng.search(q[qi:qi + 1], K, NPROBE, tiers=L.TIER_BOTH, out=out)
            if not np.array_equal(ids[0], ref[0][qi]):
                ok[0] = False
            counts[t] += 1
            i += threads
            if counts[t] % 64 == 0:
                batch_calls.append(eng.stats().last_batch_calls)

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    total = sum(counts)
    return total / dt, (float(np.mean(batch_calls)) if batch_calls else 1.0), ok[0]


for T in [int(a) for a in sys.argv[1:]] or [1, 4, 16, 64]:
    for co in (True, False):
        qps, avg_batch, ok = run(T, co)
        print(f"threads {T:3d} coalesce {int(co)}: {qps:9.0f} calls/s  (avg calls per device batch {avg_batch:5.1f}, "
              f"latency {1e3 * T / qps:6.3f} ms, results identical to the batched reference: {ok})", flush=True)
