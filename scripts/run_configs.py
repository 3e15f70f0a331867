"""BASELINE.json configs[2] and configs[3] at full size on one B200: timing + parity samples.

    python scripts/run_configs.py kmeans     # 1M x 384, nlist=4096, 20 Lloyd iterations, shared init
    python scripts/run_configs.py filtered   # 300K recent + 700K IVF, 10 % filter bitmap, 1 % tombstones

Prints one JSON line per config (kept under profiles/).  Parity is checked against the CPU oracle
on a bounded sample (the oracle costs ~0.5 us per 384-d distance)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle as O  # noqa: E402
from fabstir_vectordb_b200 import Engine, _lib as L, synth  # noqa: E402

DIM, SEED, SEED_Q, SIGMA = 384, 1234, 5678, 1.0
torch.cuda.set_device(0)
lib = L.load()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream


def gen_rows(r0, n, n_comp):
    buf = torch.empty((n, DIM), dtype=torch.float32, device=dev)
    assert lib.fvdb_synth_rows_device(buf.data_ptr(), r0, n, DIM, n_comp, SIGMA, SEED, stream) == 0
    torch.cuda.synchronize()
    return buf


def run_kmeans():
    n, nlist, iters = int(os.environ.get("FVDB_KM_ROWS", 1_000_000)), int(os.environ.get("FVDB_KM_NLIST", 4096)), 20
    n_comp = 4 * nlist
    data = gen_rows(0, n, n_comp)
    init = data[torch.arange(nlist, device=dev) * (n // nlist)].contiguous()
    eng = Engine(DIM, k_max=16)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = eng.train_device(data.data_ptr(), n, nlist, iters, init.data_ptr(), SEED)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    cents = eng.get_centroids()
    # parity sample: the assignment of 2048 rows against the FINAL centroids must equal the oracle's
    # find_nearest_centroid (src/ivf/core.rs:373-386); plus one Lloyd step on a 20K-row sample
    rows = np.linspace(0, n - 1, 2048).astype(np.int64)
    xs = data[torch.from_numpy(rows).to(dev)].cpu().numpy()
    a_gpu = eng.assign(xs)
    a_cpu = O.assign(xs, cents)
    flops = 2.0 * n * nlist * DIM * res["iterations"]
    line = {"config": f"k-means {n}x{DIM} nlist={nlist} max_iter={iters} (shared init: every {n // nlist}-th row)",
            "iterations": res["iterations"], "converged": res["converged"],
            "initial_error": res["initial_error"], "final_error": res["final_error"],
            "seconds_total": dt, "seconds_per_iteration": dt / max(1, res["iterations"]),
            "assignment_tflops": flops / dt / 1e12,
            "parity_sample": {"rows": int(len(rows)), "identical_assignments": int((a_gpu == a_cpu).sum())}}
    print(json.dumps(line), flush=True)
    eng.close()


def run_filtered():
    n_total, n_recent, nlist, nq, k, nprobe = 1_000_000, 300_000, 1024, 1024, 10, 32
    n_comp = 4 * nlist
    eng = Engine(DIM, k_max=16)
    log = lambda m: print(m, file=sys.stderr, flush=True)
    # centroids: same recipe as bench.py; rows [0, 700K) -> IVF, [700K, 1M) -> recent tier
    n_train = 64 * nlist
    train = torch.empty((n_train, DIM), dtype=torch.float32, device=dev)
    assert lib.fvdb_synth_rows_strided_device(train.data_ptr(), 0, n_train, DIM, n_comp, SIGMA, SEED, 64,
                                              max(1, n_total // n_train), stream) == 0
    torch.cuda.synchronize()
    init = train[torch.arange(nlist, device=dev) * (n_train // nlist)].contiguous()
    eng.train_device(train.data_ptr(), n_train, nlist, 8, init.data_ptr(), SEED)
    CH = 1 << 18
    n_ivf = n_total - n_recent
    for r0 in range(0, n_total, CH):
        n = min(CH, n_total - r0)
        buf = gen_rows(r0, n, n_comp)
        ids = torch.arange(r0, r0 + n, dtype=torch.int32, device=dev)
        a = max(0, min(n, n_ivf - r0))
        if a > 0:
            eng.ivf_add_device(buf.data_ptr(), ids.data_ptr(), a)
        if a < n:
            eng.flat_add_device(buf[a:].data_ptr(), ids[a:].data_ptr(), n - a)
    fbits = synth.filter_bitmap((n_total + 63) // 64 * 64, 10, 99)          # 10 % of the rows pass
    dele = np.arange(7, n_total, 100, dtype=np.uint32)                        # 1 % tombstones
    eng.set_deleted(dele, True)
    q = bench.make_queries(torch, lib, nq, n_total, n_comp, 0)
    d_f = torch.from_numpy(fbits.view(np.int64)).to(dev)
    o_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    o_dst = torch.empty((nq, k), dtype=torch.float32, device=dev)
    o_cnt = torch.empty((nq,), dtype=torch.int32, device=dev)

    def step():
        eng.search_device(q.data_ptr(), nq, k, nprobe, L.TIER_BOTH, d_f.data_ptr(), fbits.size * 64,
                          o_ids.data_ptr(), o_dst.data_ptr(), o_cnt.data_ptr(), stream)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    steps = 10
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    st = eng.stats()
    # parity on a sample of queries against the oracle (pre-filter semantics + tombstones)
    ns = 24
    x_host = np.empty((n_total, DIM), dtype=np.float32)
    for r0 in range(0, n_total, CH):
        n = min(CH, n_total - r0)
        x_host[r0:r0 + n] = gen_rows(r0, n, n_comp).cpu().numpy()
    cents = eng.get_centroids()
    assign = eng.assign(x_host[:n_ivf])
    ivf = O.IVF(cents, x_host[:n_ivf], np.arange(n_ivf, dtype=np.uint32), assign_=assign)
    qs = q[:ns].cpu().numpy()
    want = O.hybrid_batch_search(ivf, x_host[n_ivf:], np.arange(n_ivf, n_total, dtype=np.uint32), qs, k, nprobe,
                                 tiers=3, deleted=O.make_bitmap(n_total, dele), filter_bits=fbits)
    g_ids = o_ids[:ns].cpu().numpy().view(np.uint32)
    g_dst = o_dst[:ns].cpu().numpy()
    same_ids = int((g_ids == want[0]).all(axis=1).sum())
    same_bits = int((g_dst.view(np.uint32) == want[1].view(np.uint32)).all(axis=1).sum())
    line = {"config": "hybrid filtered search 300K recent + 700K IVF, nlist=1024 nprobe=32 nq=1024 k=10, "
                      "filter bitmap 10 % pass, 1 % tombstones",
            "ms_per_batch": ms, "qps": nq / (ms * 1e-3), "scan_ms": st.last_scan_ms,
            "fallback_queries": int(st.last_fallback_queries), "launches": int(st.last_launches),
            "parity_sample": {"queries": ns, "identical_id_lists": same_ids, "identical_distance_bits": same_bits}}
    print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("kmeans", "all"):
        run_kmeans()
    if which in ("filtered", "all"):
        run_filtered()
