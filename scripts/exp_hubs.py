"""Probe-count statistics of the bench index: what the hub lists (kernel W's share) cost in row passes
under different item-shaping rules (host replay of the bucketing).  python scripts/exp_hubs.py"""
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
n_total, nlist, n_comp = bench.build_index(torch, eng, 0, 1, lambda m: None)
q = bench.make_queries(torch, lib, bench.NQ_PER_GPU, n_total, n_comp, 0)
keys = torch.empty((q.shape[0], bench.NPROBE), dtype=torch.int64, device="cuda")
eng.coarse_device(q.data_ptr(), q.shape[0], bench.NPROBE, keys.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
lists = (keys.cpu().numpy().view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64)
cnt = np.bincount(lists.ravel(), minlength=nlist)
ids, li = eng.dump_lists() if False else (None, None)
x = None
a = torch.empty((1 << 18,), dtype=torch.int32, device="cuda")
sizes = np.zeros(nlist, dtype=np.int64)
buf = torch.empty((1 << 18, bench.DIM), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for r0 in range(0, n_total, 1 << 18):
    n = min(1 << 18, n_total - r0)
    lib.fvdb_synth_rows_device(buf.data_ptr(), r0, n, bench.DIM, n_comp, bench.SIGMA, bench.SEED, st)
    eng.assign_device(buf.data_ptr(), n, a.data_ptr(), st)
    torch.cuda.synchronize()
    sizes += np.bincount(a[:n].cpu().numpy(), minlength=nlist)
print("lists", nlist, "rows", sizes.sum(), "pairs(query,row)", int((cnt * sizes).sum()))
for wm in (65, 129, 193, 257):
    hub = cnt >= wm
    print(f"wide_min {wm}: hub lists {hub.sum()} rows {sizes[hub].sum()} pairs {(cnt*sizes)[hub].sum()/1e6:.1f}M | "
          f"R row-passes {(np.ceil(cnt[~hub]/64)*sizes[~hub]).sum()/1e3:.0f}K  W row-passes {(np.ceil(cnt[hub]/256)*sizes[hub]).sum()/1e3:.0f}K")
hub = cnt >= 129
c, s = cnt[hub], sizes[hub]
for tail in (0, 32, 64, 128):
    r = c % 256
    to_r = (r > 0) & (r <= tail)
    w_pass = (np.where(to_r, c // 256, np.ceil(c / 256)) * s).sum()
    r_pass = (np.where(to_r, np.ceil(r / 64), 0) * s).sum()
    print(f"tail<= {tail}: W row-passes {w_pass/1e3:.0f}K (MMA cost ~{w_pass*114/148/1.9e3:.0f} us at 57 cyc/row/pair) + R row-passes {r_pass/1e3:.0f}K (~{r_pass*16.5/148/1.9e3:.0f} us)")
print("hub probe counts:", np.sort(c)[::-1][:40].tolist())
print("hub sizes       :", s[np.argsort(-c)][:40].tolist())
