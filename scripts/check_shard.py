"""Multi-GPU parity check (run under torchrun, one rank per GPU): list-sharded search + NCCL
all-gather + merge must equal the unsharded CPU oracle on EVERY rank, bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 scripts/check_shard.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from fabstir_vectordb_b200 import Engine, _lib as L, synth  # noqa: E402
from fabstir_vectordb_b200.shard import ShardedIndex  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
N, D, NLIST, NQ, K, NPROBE = 40_000, 384, 64, 96, 10, 8
x = synth.rows(0, N, D, 4 * NLIST, 1.0, 1234)
q = synth.queries(0, NQ, D, N, 4 * NLIST, 1.0, 1234, synth.default_qnoise(D, 1.0), 5678)
cents = x[:: N // NLIST][:NLIST].copy()
ok = True
for mode in (L.SCAN_EXACT, L.SCAN_TC):
    eng = Engine(D, k_max=16, device=local)
    eng.set_option(L.OPT_SCAN_MODE, mode)
    eng.set_centroids(cents)
    sh = ShardedIndex(eng, rank, world)
    dx = torch.from_numpy(x).to(dev)
    ids = torch.arange(N, dtype=torch.int32, device=dev)
    kept = sh.add_rows_device(dx, ids)
    dq = torch.from_numpy(q).to(dev)
    g_ids, g_dst, g_cnt = sh.search(dq, K, NPROBE, tiers=L.TIER_HISTORICAL)
    torch.cuda.synchronize()
    ivf = O.IVF(cents, x, np.arange(N, dtype=np.uint32))
    w_ids, w_dst, w_cnt = O.hybrid_batch_search(ivf, None, None, q, K, NPROBE, tiers=2)
    a = g_ids.cpu().numpy().view(np.uint32)
    same = (a == w_ids).all() and (g_dst.cpu().numpy().view(np.uint32) == w_dst.view(np.uint32)).all() \
        and (g_cnt.cpu().numpy().view(np.uint32) == w_cnt).all()
    print(f"rank {rank}/{world} mode {mode}: kept {kept} rows, global result identical to the oracle: {bool(same)}", flush=True)
    ok = ok and bool(same)
    eng.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
