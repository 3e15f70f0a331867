"""Generates tests/golden/*.npz: seeded inputs (the counter-based generator of
fabstir_vectordb_b200/synth.py) and the CPU oracle's outputs for them.  The GPU tests compare the
CUDA path with these files WITHOUT running the oracle; a CPU test checks that the oracle still
reproduces them (so a change to either side shows up).  The Rust reference cannot run here, so the
oracle — pinned by the reference's known-answer tests (reference_kats.json) — is the generator.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402
from fabstir_vectordb_b200 import synth  # noqa: E402

CASES = {
    # name: (n_ivf, n_flat, dim, nlist, nq, k, nprobe)
    "hybrid_d384": (6000, 900, 384, 24, 40, 10, 6),
    "hybrid_d64": (3000, 300, 64, 16, 33, 7, 4),
}


def case_inputs(name):
    n, nf, d, nlist, nq, k, nprobe = CASES[name]
    x = synth.rows(0, n + nf, d, 4 * nlist, 0.8, 4242)
    q = synth.queries(0, nq, d, n + nf, 4 * nlist, 0.8, 4242, synth.default_qnoise(d, 0.8), 777)
    cents = x[:: n // nlist][:nlist].copy()
    deleted = np.arange(5, n + nf, 61, dtype=np.uint32)
    keep = np.arange(0, n + nf, dtype=np.uint32)
    keep = keep[((keep.astype(np.uint64) * 2654435761) % (1 << 32)) % 5 != 0]   # ~80 % of the rows pass the filter
    return x, q, cents, deleted, keep


def main():
    for name, (n, nf, d, nlist, nq, k, nprobe) in CASES.items():
        x, q, cents, deleted, keep = case_inputs(name)
        ids = np.arange(n, dtype=np.uint32)
        fid = np.arange(n, n + nf, dtype=np.uint32)
        ivf = O.IVF(cents, x[:n], ids)
        out = {}
        r = O.hybrid_batch_search(ivf, x[n:], fid, q, k, nprobe, tiers=3)
        out["plain_ids"], out["plain_dist"], out["plain_cnt"] = r
        dbits = O.make_bitmap(n + nf, deleted)
        fbits = O.make_bitmap(n + nf, keep)
        r = O.hybrid_batch_search(ivf, x[n:], fid, q, k, nprobe, tiers=3, deleted=dbits, filter_bits=fbits)
        out["masked_ids"], out["masked_dist"], out["masked_cnt"] = r
        out["assign"] = ivf.assign
        init = cents.copy()
        cent2, assign2, res = O.train_lloyd(x[:n], init, 4)
        out["lloyd_centroids"] = cent2
        out["lloyd_iterations"] = np.array([res["iterations"]], dtype=np.uint32)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k_: v.shape for k_, v in out.items()})


if __name__ == "__main__":
    main()
