"""numpy twin of csrc/synth.cu (include/fvdb_synth.h): counter-based Gaussian-mixture data,
L2-normalised.  Integer hashing + correctly-rounded fp32 operations only, so these arrays are
bit-identical to what the device generator writes (tests/test_synth.py checks that on a GPU)."""
from __future__ import annotations

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)
_A = np.uint64(0xD1B54A32D192ED03)
_B = np.uint64(0x8CB92BA72F3D8DD7)
_C = np.uint64(0x2545F4914F6CDD1D)
_GS = np.float32(2.6429157e-05)


def _mix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(30))
        x = x * _M1
        x = x ^ (x >> np.uint64(27))
        x = x * _M2
        x = x ^ (x >> np.uint64(31))
    return x


def _stream_key(seed: int, stream: int):
    with np.errstate(over="ignore"):
        return _mix64(np.uint64(seed) + np.uint64(stream) * _G)


def _hash2(key, a, b):
    with np.errstate(over="ignore"):
        return _mix64(key ^ (np.asarray(a, np.uint64) * _A + np.asarray(b, np.uint64) * _B + _C))


def _gauss(h):
    s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)) +
         ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + (h >> np.uint64(48))).astype(np.int64)
    return (s - 131070).astype(np.float32) * _GS


def _raw_rows(rows, dim, n_comp, sigma, seed):
    rows = np.asarray(rows, dtype=np.uint64)
    kc, kx = _stream_key(seed, 1), _stream_key(seed, 2)
    d = np.arange(dim, dtype=np.uint64)[None, :]
    comp = (rows % np.uint64(n_comp))[:, None]
    c = _gauss(_hash2(kc, comp, d))
    e = _gauss(_hash2(kx, rows[:, None], d))
    return c + np.float32(sigma) * e  # fp32 mul then fp32 add


def _normalise(v):
    sq = v * v
    ss = np.cumsum(sq, axis=1, dtype=np.float32)[:, -1]  # sequential fp32 sum
    inv = np.float32(1.0) / np.sqrt(ss)
    return v * inv[:, None]


def rows(row0: int, n: int, dim: int, n_comp: int, sigma: float, seed: int, chunk: int = 8192):
    out = np.empty((n, dim), dtype=np.float32)
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        r = np.arange(row0 + b, row0 + e, dtype=np.uint64)
        out[b:e] = _normalise(_raw_rows(r, dim, n_comp, sigma, seed))
    return out


def query_base_rows(q0: int, n: int, n_total: int, seed_q: int):
    kb = _stream_key(seed_q, 3)
    qi = np.arange(q0, q0 + n, dtype=np.uint64)
    return _hash2(kb, qi, np.uint64(0)) % np.uint64(n_total)


def queries(q0: int, n: int, dim: int, n_total: int, n_comp: int, sigma: float, seed: int,
            qnoise: float, seed_q: int):
    base = query_base_rows(q0, n, n_total, seed_q)
    x = _normalise(_raw_rows(base, dim, n_comp, sigma, seed))
    kn = _stream_key(seed_q, 4)
    qi = np.arange(q0, q0 + n, dtype=np.uint64)[:, None]
    d = np.arange(dim, dtype=np.uint64)[None, :]
    v = x + np.float32(qnoise) * _gauss(_hash2(kn, qi, d))
    return _normalise(v)


def filter_bitmap(nbits: int, mod: int, seed: int):
    assert nbits % 64 == 0
    kf = _stream_key(seed, 5)
    ids = np.arange(nbits, dtype=np.uint64)
    hit = (_hash2(kf, ids, np.uint64(0)) % np.uint64(mod)) == 0
    bits = hit.reshape(-1, 64).astype(np.uint64) << np.arange(64, dtype=np.uint64)[None, :]
    return np.bitwise_or.reduce(bits, axis=1)


def default_qnoise(dim: int, sigma: float) -> float:
    """0.1 x the per-dimension noise std of a normalised row."""
    return float(0.1 * sigma / np.sqrt(dim * (1.0 + sigma * sigma)))
