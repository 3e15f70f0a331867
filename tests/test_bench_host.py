"""CPU checks of bench.py's host-side logic (no GPU, no timing): the training-sample layout both arms share, the
k-means seeding rule, the reference arm's host-built index on a tiny workload, and that the committed scan-traffic
stamp belongs to the committed kernel sources (otherwise the bench line would carry no `roofline.traffic`)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle as O  # noqa: E402
from fabstir_vectordb_b200 import synth  # noqa: E402


def test_training_sample_is_blocks_of_64_rows():
    rows = bench.train_sample_rows(1_000_000, 65536)
    assert rows.shape == (65536,)
    stride = 1_000_000 // 65536
    assert rows[:64].tolist() == list(range(64))
    assert rows[64] == 64 * stride and rows[65] == 64 * stride + 1
    assert rows.max() < 1_000_000 and len(set(rows.tolist())) == 65536


def test_kmeans_seeds_come_from_distinct_mixture_components():
    nlist, n_train, n_total = 1024, 65536, 1_000_000
    seeds = bench.train_sample_rows(n_total, n_train)[bench.init_rows_of_sample(nlist, n_train)]
    comps = seeds % np.uint64(bench.n_comp_for(nlist))
    assert len(set(comps.tolist())) == nlist          # the clumped seeding of early round 1 gave nlist / 16


def test_nprobe_and_workload_names():
    assert bench.nprobe_for(1) == 32 and bench.nprobe_for(8) == 32
    assert "nlist=8192" in bench.workload_name(8) and "8000000x384" in bench.workload_name(8)


def test_reference_arm_host_index_small(monkeypatch):
    """host_index (numpy generator + the oracle's k-means and assignment, no product library) on a tiny workload:
    the lists partition the rows, every row sits in the list of its nearest centroid, and a search finds a
    database row from a slightly perturbed copy of it."""
    monkeypatch.setattr(bench, "TRAIN_ITERS", 4)
    n_total, nlist = 4096, 16
    n_comp = bench.n_comp_for(nlist)
    ivf, x = bench.host_index(n_total, nlist, n_comp, lambda m: None)
    assert x.shape == (n_total, bench.DIM)
    assert np.array_equal(x[:8], synth.rows(0, 8, bench.DIM, n_comp, bench.SIGMA, bench.SEED))
    assert sum(ivf.list_len(l) for l in range(nlist)) == n_total
    assert np.array_equal(ivf.assign, O.assign(x, ivf.centroids))
    q = bench.host_queries(8, n_total, n_comp, 0)
    ids, dist, cnt = O.hybrid_batch_search(ivf, None, None, q, 5, nlist, tiers=2)
    base = synth.query_base_rows(0, 8, n_total, bench.SEED_Q)
    assert (cnt == 5).all() and ids[:, 0].tolist() == base.tolist()


def test_committed_scan_traffic_stamp_matches_the_kernel_sources():
    with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as fh:
        tj = json.load(fh)
    assert tj["source_digest"] == bench.scan_source_digest(), \
        "the scan kernels changed after the ncu capture: re-capture (scripts/gpu_evidence.sh + summarize_ncu.py --traffic)"
    assert 1.0e9 < tj["dram_bytes_per_launch"] < 2.0e9
