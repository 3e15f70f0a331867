"""One rank's share of an N-GPU sharded search, reproduced on ONE GPU (no NCCL), so that it can be profiled
with ncu: the index is rank 0's shard of the N-times larger index, the batch the N-times larger batch; the
coarse ranking of the whole batch is computed locally (what the all-gather would deliver).

    python scripts/exp_rank_of.py 8 [steps]
"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from fabstir_vectordb_b200 import Engine, _lib as L
from fabstir_vectordb_b200.shard import ShardedIndex, place_lists

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.cuda.set_device(0)
lib = L.load()
eng = Engine(bench.DIM, k_max=16)
sh = ShardedIndex(eng, 0, world)          # world > 1 only for the loader; no collective is called below
log = lambda m: print(m, file=sys.stderr, flush=True)
n_total, nlist, n_comp = bench.build_index(torch, eng, 0, world, log, sh=sh)
nq = bench.NQ_PER_GPU * world
NP = bench.nprobe_for(world)
qs = [bench.make_queries(torch, lib, nq, n_total, n_comp, s) for s in range(4)]
stream = torch.cuda.current_stream().cuda_stream
keys = torch.empty((nq, NP), dtype=torch.int64, device="cuda")
out = (torch.empty((nq, bench.K), dtype=torch.int32, device="cuda"), torch.empty((nq, bench.K), dtype=torch.float32, device="cuda"),
       torch.empty((nq,), dtype=torch.int32, device="cuda"))
per = nq // world
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
acc = np.zeros(3)
scan = []
if os.environ.get("FVDB_BENCH_PROFILE"):
    torch.cuda.profiler.start()
for s in range(steps + 3):
    q = qs[s % 4]
    eng.coarse_device(q.data_ptr(), nq, NP, keys.data_ptr(), stream)      # the whole ranking (not timed as such)
    ev[0].record()
    eng.coarse_device(q[:per].data_ptr(), per, NP, keys[:per].data_ptr(), stream)   # this rank's slice
    ev[1].record()
    eng.search_device_coarse(q.data_ptr(), nq, bench.K, NP, L.TIER_HISTORICAL, 0, 0, keys.data_ptr(),
                             out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), stream)
    ev[2].record()
    torch.cuda.synchronize()
    if s >= 3:
        acc += np.array([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), 0.0])
        st = eng.stats()
        scan.append(st.last_scan_ms)
if os.environ.get("FVDB_BENCH_PROFILE"):
    torch.cuda.profiler.stop()
acc /= steps
print(f"rank 0 of {world}: nq={nq} nprobe={NP} rows={eng.stats().ivf_rows}: coarse slice {acc[0]:.3f} ms, local search {acc[1]:.3f} ms "
      f"(scan {np.mean(scan):.3f} ms, launches {eng.stats().last_launches})", flush=True)
