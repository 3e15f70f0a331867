"""Design experiment (GPU, torch): how many probed rows fall below candidate seed thresholds,
neighbour-distance gaps, and the real TF32 error — on the bench index (1M x 384, nlist 1024)."""
import os, sys, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from fabstir_vectordb_b200 import Engine, _lib as L

torch.cuda.set_device(0)
lib = L.load()
dev = torch.device("cuda")
N, D, NL, NP = bench.ROWS_PER_GPU, bench.DIM, bench.NLIST_PER_GPU, bench.NPROBE
n_comp = bench.n_comp_for(NL)
eng = Engine(D, k_max=16)
eng.set_option(L.OPT_SCAN_MODE, L.SCAN_EXACT)
log = lambda m: print(m, file=sys.stderr)
bench.build_index(torch, eng, 0, 1, log)
cents = torch.from_numpy(eng.get_centroids()).to(dev)
X = torch.empty((N, D), dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
CH = 1 << 18
for r0 in range(0, N, CH):
    n = min(CH, N - r0)
    assert lib.fvdb_synth_rows_device(X[r0:].data_ptr(), r0, n, D, n_comp, bench.SIGMA, bench.SEED, st) == 0
torch.cuda.synchronize()
torch.backends.cuda.matmul.allow_tf32 = False
xn = (X * X).sum(1)
cn = (cents * cents).sum(1)
assign = torch.empty(N, dtype=torch.long, device=dev)
for r0 in range(0, N, CH):
    d = xn[r0:r0+CH, None] + cn[None, :] - 2 * X[r0:r0+CH] @ cents.T
    assign[r0:r0+CH] = d.argmin(1)
sizes = torch.bincount(assign, minlength=NL)
print("list sizes: min %d med %d max %d" % (sizes.min(), sizes.median(), sizes.max()))
NQ = 256
q = bench.make_queries(torch, lib, 1024, N, n_comp, 0)[:NQ]
qn = (q * q).sum(1)
dc = qn[:, None] + cn[None, :] - 2 * q @ cents.T
probe = dc.topk(NP, largest=False).indices          # [NQ, NP]
d2 = (qn[:, None] + xn[None, :] - 2 * (q.double() @ X.double().T).float()).clamp_min(0)   # exact-ish
torch.backends.cuda.matmul.allow_tf32 = True
d2_tf = (qn[:, None] + xn[None, :] - 2 * (q @ X.T)).clamp_min(0)  # cuBLAS TF32 (round-to-nearest inputs)
torch.backends.cuda.matmul.allow_tf32 = False
err = (d2_tf - d2).abs()
res = {"tf32_abs_err_d2": {"mean": err.mean().item(), "p99": err.flatten()[::97].quantile(0.99).item(), "max": err.max().item()}}
inprobe = torch.zeros((NQ, NL), dtype=torch.bool, device=dev)
inprobe.scatter_(1, probe, True)
mask = inprobe[:, assign]                            # [NQ, N] row is in a probed list
d2p = torch.where(mask, d2, torch.full_like(d2, float("inf")))
srt = d2p.sort(1).values
g = {}
for a, b in ((10, 16), (10, 20), (10, 24), (10, 32), (10, 48), (10, 64)):
    gap = (srt[:, b - 1] - srt[:, a - 1])
    g[f"gap_d2_{a}_{b}"] = {"min": gap.min().item(), "p05": gap.quantile(0.05).item(), "med": gap.median().item()}
res["d2_10th"] = {"med": srt[:, 9].median().item(), "min": srt[:, 9].min().item()}
res["gaps"] = g
# seed thresholds
nearest = probe[:, 0]
counts = {}
first_row = torch.zeros(NL, dtype=torch.long, device=dev)
order = torch.argsort(assign, stable=True)
offs = torch.cumsum(sizes, 0) - sizes
for name, take in (("first64", 64), ("first128", 128), ("first256", 256), ("whole", 10**9)):
    cs = []
    for i in range(NQ):
        l = nearest[i].item()
        rows = order[offs[l]: offs[l] + min(take, sizes[l].item())]
        dd = d2[i, rows].sort().values
        kth = dd[min(31, len(dd) - 1)]
        cs.append((d2p[i] < kth).sum().item())
    cs = np.array(cs)
    counts[name + "_32nd"] = {"med": float(np.median(cs)), "p95": float(np.quantile(cs, 0.95)), "max": int(cs.max())}
res["candidates_below_seed"] = counts
# rank of the list that holds each of the true top-10 (how concentrated in the nearest lists)
top10 = d2p.topk(10, largest=False).indices
l10 = assign[top10]                                   # [NQ,10]
rank_of = torch.full((NQ, NL), NP, dtype=torch.long, device=dev)
rank_of.scatter_(1, probe, torch.arange(NP, device=dev)[None, :].expand(NQ, NP))
rk = rank_of.gather(1, l10)
res["top10_list_rank"] = {"frac_rank0": (rk == 0).float().mean().item(), "frac_rank<4": (rk < 4).float().mean().item(), "mean": rk.float().mean().item()}
print(json.dumps(res, indent=1))
