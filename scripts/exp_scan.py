"""Scan-kernel pipeline experiments on the bench index (1M x 384, nlist 1024, nprobe 32, 1024
queries): runs the TC search under a list of FVDB_TC_DEBUG values (timing-only modes that switch
off parts of the pipeline, see tc_scan.cu) and prints the device time of the scan kernel.

    python scripts/exp_scan.py 0 5 37 ...        # debug values; optional KEY=VALUE env pairs
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fabstir_vectordb_b200 import Engine, _lib as L  # noqa: E402
from fabstir_vectordb_b200.shard import ShardedIndex  # noqa: E402

torch.cuda.set_device(0)
lib = L.load()
eng = Engine(bench.DIM, k_max=16)
eng.set_option(L.OPT_SCAN_MODE, L.SCAN_TC)
log = lambda m: print(m, file=sys.stderr, flush=True)
n_total, nlist, n_comp = bench.build_index(torch, eng, 0, 1, log)
sh = ShardedIndex(eng, 0, 1)
qsets = [bench.make_queries(torch, lib, bench.NQ_PER_GPU, n_total, n_comp, s) for s in range(4)]
sh.search(qsets[0], bench.K, bench.NPROBE, tiers=L.TIER_HISTORICAL)
torch.cuda.synchronize()

for arg in sys.argv[1:]:
    envs = {}
    parts = arg.split(",")
    dbg = parts[0]
    for kv in parts[1:]:
        k, v = kv.split("=")
        envs[k] = v
    os.environ["FVDB_TC_DEBUG"] = dbg
    for k, v in envs.items():
        os.environ[k] = v
    ms, tot = [], []
    try:
        for i in range(8):
            sh.search(qsets[0 if "ONESET" in envs else i % 4], bench.K, bench.NPROBE, tiers=L.TIER_HISTORICAL)
            torch.cuda.synchronize()
            st = eng.stats()
            if i >= 3:
                ms.append(st.last_scan_ms)
                tot.append(st.last_device_ms)
        print(f"debug {dbg:>6s} {envs} scan_ms {np.mean(ms):.4f} (min {np.min(ms):.4f}) device_ms {np.mean(tot):.4f} "
              f"fallback {st.last_fallback_queries}", flush=True)
    except Exception as e:  # a debug mode that faults ends the experiment
        print(f"debug {dbg} FAILED: {e}", flush=True)
        break
    for k in envs:
        os.environ.pop(k, None)
