"""How many queries probe each list (bench workload)?  Items per list at several query-tile sizes."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from fabstir_vectordb_b200 import Engine, _lib as L
torch.cuda.set_device(0)
lib = L.load()
eng = Engine(bench.DIM, k_max=16)
log = lambda m: print(m, file=sys.stderr, flush=True)
n_total, nlist, n_comp = bench.build_index(torch, eng, 0, 1, log)
cents = torch.from_numpy(eng.get_centroids()).cuda()
q = bench.make_queries(torch, lib, bench.NQ_PER_GPU, n_total, n_comp, 0)
d = (q * q).sum(1)[:, None] + (cents * cents).sum(1)[None, :] - 2 * q @ cents.T
probe = d.topk(bench.NPROBE, largest=False).indices
cnt = torch.bincount(probe.flatten(), minlength=nlist).cpu().numpy()
x = np.empty((n_total,), dtype=np.int64)
# list sizes via assign of all rows is expensive; use the engine's view: sizes from a search of stats are not exposed -> approximate with equal sizes
print("pairs per list: mean %.1f median %d p90 %d p99 %d max %d, lists with 0: %d" % (cnt.mean(), np.median(cnt), np.percentile(cnt, 90), np.percentile(cnt, 99), cnt.max(), (cnt == 0).sum()))
for tq in (32, 64, 96, 128, 192, 256):
    items = np.ceil(cnt / tq).sum()
    print(f"TQ={tq}: items {int(items)}  restream factor {items / (cnt > 0).sum():.2f}")
