"""Per-phase device timing of the sharded search (torchrun, one rank per GPU): coarse slice,
coarse all-gather, local search, result all-gathers, merge.  CUDA events on the current stream.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/prof_shard.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fabstir_vectordb_b200 import Engine, _lib as L  # noqa: E402
from fabstir_vectordb_b200.shard import ShardedIndex  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = L.load()
eng = Engine(bench.DIM, k_max=16, device=local)
log = lambda m: None
sh = ShardedIndex(eng, rank, world)
n_total, nlist, n_comp = bench.build_index(torch, eng, rank, world, log, sh=sh)
nq = bench.NQ_PER_GPU * world
qsets = [bench.make_queries(torch, lib, nq, n_total, n_comp, s) for s in range(4)]
K, NP = bench.K, bench.nprobe_for(world)
for i in range(6):
    sh.search(qsets[i % 4], K, NP)
torch.cuda.synchronize()

names = ["coarse_slice", "coarse_allgather", "local_search", "result_allgather", "merge"]
acc = np.zeros(len(names))
scan = []
steps = 12
stream = torch.cuda.current_stream().cuda_stream
for s in range(steps):
    q = qsets[s % 4]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    b = sh._buffers(nq, K, q.device)
    np_ = min(NP, eng.stats().nlist)
    per = (nq + world - 1) // world
    mine, allk = sh._bufs[("coarse", nq, np_)] if world > 1 else (None, None)
    lo = rank * per
    ev[0].record()
    if world > 1:
        if sh.share_bounds:
            eng.bounds_begin_batch(nq, stream)
        eng.coarse_device(q[lo:lo + per].data_ptr(), per, np_, mine.data_ptr(), stream)
    ev[1].record()
    if world > 1:
        dist.all_gather_into_tensor(allk, mine)
    ev[2].record()
    if world > 1:
        eng.search_device_coarse(q.data_ptr(), nq, K, np_, L.TIER_HISTORICAL, 0, 0, allk.data_ptr(),
                                 b["ids"].data_ptr(), b["dist"].data_ptr(), b["cnt"].data_ptr(), stream)
    else:
        eng.search_device(q.data_ptr(), nq, K, np_, L.TIER_HISTORICAL, 0, 0, b["ids"].data_ptr(),
                          b["dist"].data_ptr(), b["cnt"].data_ptr(), stream)
    scan.append(eng.stats().last_scan_ms)
    ev[3].record()
    if world > 1:
        dist.all_gather_into_tensor(b["g_pack"].view(-1), b["pack"])
    ev[4].record()
    if world > 1:
        eng.merge_topk_packed_device(b["g_pack"].data_ptr(), world, nq, K, b["o_ids"].data_ptr(),
                                     b["o_dist"].data_ptr(), b["o_cnt"].data_ptr(), stream)
    ev[5].record()
    torch.cuda.synchronize()
    acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))])
acc /= steps
print(f"rank {rank}/{world} nq={nq}: " + "  ".join(f"{n}={v:.3f}" for n, v in zip(names, acc)) +
      f"  total={acc.sum():.3f} ms  (scan kernel {np.mean(scan):.3f} ms)", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
