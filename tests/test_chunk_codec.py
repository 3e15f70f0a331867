"""VectorChunk CBOR codec (include/fvdb_chunk.h): the native codec against the
pure-Python oracle (oracle/cbor_chunk.py), the RFC 8949 Appendix A item encodings as known answers,
and the reference's own chunk tests (round trips: tests/unit/chunk_tests.rs:38-70, :340-360).

Host functions only: runs without a GPU.  Byte-level parity with serde_cbor itself is UNPINNED (the
reference holds no CBOR byte fixture and serde_cbor cannot run here); what is pinned: the item layer
by RFC 8949's published examples, and both directions of the codec by an independent restatement.
"""
import ctypes as C
import struct

import numpy as np
import pytest

from fabstir_vectordb_b200 import _lib as L
from fabstir_vectordb_b200.chunk import ChunkError, decode_vector_chunk, encode_vector_chunk
from fabstir_vectordb_b200.engine import InconsistentDimensions
from oracle import cbor_chunk as O


def _ids(n, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (n, 32), dtype=np.uint8)


def _same_bits(a, b):
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


# ---- known answers: RFC 8949 Appendix A ----------------------------------------------------------
RFC_UINTS = [(0, "00"), (1, "01"), (10, "0a"), (23, "17"), (24, "1818"), (25, "1819"), (100, "1864"),
             (1000, "1903e8"), (1000000, "1a000f4240"), (1000000000000, "1b000000e8d4a51000"),
             (18446744073709551615, "1bffffffffffffffff")]
RFC_FLOATS = [(0.0, "f90000"), (-0.0, "f98000"), (1.0, "f93c00"), (1.5, "f93e00"), (65504.0, "f97bff"),
              (100000.0, "fa47c35000"), (3.4028234663852886e+38, "fa7f7fffff"),
              (5.960464477539063e-8, "f90001"), (0.00006103515625, "f90400"), (-4.0, "f9c400"),
              (float("inf"), "f97c00"), (float("nan"), "f97e00"), (float("-inf"), "f9fc00")]


@pytest.mark.parametrize("value,hexs", RFC_UINTS)
def test_rfc8949_unsigned_integers(value, hexs):
    assert O.encode_uint(value).hex() == hexs
    assert O.decode_item(bytes.fromhex(hexs)) == value
    # native: start_idx carries the integer
    data = encode_vector_chunk("c", value, 0, np.zeros((0, 32), np.uint8), np.zeros((0, 0), np.float32))
    assert data == O.encode_chunk("c", value, 0, np.zeros((0, 32), np.uint8), np.zeros((0, 0), np.float32))
    assert hexs in data.hex()
    assert decode_vector_chunk(data).start_idx == value


@pytest.mark.parametrize("value,hexs", RFC_FLOATS)
def test_rfc8949_floats(value, hexs):
    assert O.encode_f32(value).hex() == hexs
    got = O.decode_item(bytes.fromhex(hexs))
    assert (np.isnan(got) and np.isnan(value)) or np.float32(got) == np.float32(value)
    # native: a one-element vector carries the float
    rows = np.array([[value]], dtype=np.float32)
    data = encode_vector_chunk("c", 0, 0, _ids(1), rows)
    assert data == O.encode_chunk("c", 0, 0, _ids(1), rows)
    assert data.hex().endswith("81" + hexs)
    back = decode_vector_chunk(data).rows
    assert (np.isnan(back[0, 0]) and np.isnan(value)) or _same_bits(back, rows)


def test_rfc8949_other_float_widths_are_read():
    # RFC 8949 App. A: 1.1 as a double, 1.0e+300 (saturates to inf in f32), -4.1 as a double, ints as numbers
    head = O.encode_chunk("c", 0, 0, _ids(1), np.zeros((1, 0), np.float32))[:-1]   # ... up to the vector's array head
    body = bytes.fromhex("85" "fb3ff199999999999a" "fbc010666666666666" "f93c00" "1864" "29")
    ch = decode_vector_chunk(head + body)
    want = np.array([[np.float32(1.1), np.float32(-4.1), 1.0, 100.0, -10.0]], dtype=np.float32)
    assert _same_bits(ch.rows, want)
    assert _same_bits(O.decode_chunk(head + body)[4], want)


# ---- the reference's own tests (round trips) ----------------------------------------------------
def test_vector_chunk_cbor_serialization():
    """tests/unit/chunk_tests.rs:38-59 — ten vectors `[i * 0.1; 4]`."""
    ids = _ids(10, 3)
    rows = np.array([[np.float32(i) * np.float32(0.1)] * 4 for i in range(10)], dtype=np.float32)
    data = encode_vector_chunk("chunk-0", 0, 9999, ids, rows)
    ch = decode_vector_chunk(data)
    assert (ch.chunk_id, ch.start_idx, ch.end_idx, len(ch)) == ("chunk-0", 0, 9999, 10)
    assert np.array_equal(ch.ids, ids) and _same_bits(ch.rows, rows)
    assert data == O.encode_chunk("chunk-0", 0, 9999, ids, rows)


def test_vector_chunk_empty_cbor_serialization():
    """tests/unit/chunk_tests.rs:61-70."""
    data = encode_vector_chunk("chunk-empty", 0, 0, np.zeros((0, 32), np.uint8), np.zeros((0, 0), np.float32))
    ch = decode_vector_chunk(data)
    assert len(ch) == 0 and ch.chunk_id == "chunk-empty" and ch.rows.shape == (0, 0)
    assert O.decode_chunk(data)[3].shape == (0, 32)


def test_large_chunk_round_trip_both_ways():
    """tests/unit/chunk_tests.rs:340-360 (10 K vectors per chunk, src/hybrid/persistence.rs:189), at
    the all-MiniLM shape; values mix f16-exact numbers (0, +-1, 0.5) with ordinary floats."""
    rng = np.random.default_rng(7)
    n, d = 2000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[rng.random((n, d)) < 0.05] = 0.0
    rows[rng.random((n, d)) < 0.02] = -1.0
    rows[rng.random((n, d)) < 0.02] = 0.5
    ids = _ids(n, 11)
    data = encode_vector_chunk("chunk-7", 70000, 79999, ids, rows)
    assert data == O.encode_chunk("chunk-7", 70000, 79999, ids, rows)       # byte for byte
    ch = decode_vector_chunk(data, pinned=False)
    assert np.array_equal(ch.ids, ids) and _same_bits(ch.rows, rows)
    o = O.decode_chunk(data)
    assert np.array_equal(o[3], ids) and _same_bits(o[4], rows)


# ---- spellings the reader must accept -------------------------------------------------------------
def test_reader_accepts_any_wellformed_spelling():
    ids = _ids(3, 5)
    rows = np.array([[1.0, 2.5, -3.25], [0.1, 0.2, 0.3], [7.0, 8.0, 9.0]], dtype=np.float32)
    T, U = O.encode_text, O.encode_uint

    def vec_indef(r):
        return b"\x9f" + b"".join(b"\xfb" + struct.pack(">d", float(x)) for x in r) + b"\xff"

    vectors = b"\xbf"                                                    # indefinite map
    vectors += b"\x58\x20" + ids[0].tobytes() + vec_indef(rows[0])         # id as a byte string, f64s
    vectors += b"\x9f" + b"".join(U(int(b)) for b in ids[1]) + b"\xff"     # id as an indefinite array
    vectors += b"\x83" + b"".join(b"\xfa" + struct.pack(">f", float(x)) for x in rows[1])
    vectors += b"\x98\x20" + b"".join(U(int(b)) for b in ids[2]) + b"\x83\x07\x08\x09"   # ints as numbers
    vectors += b"\xff"
    # fields in another order, an unknown field (skipped), the key "chunk_id" and its value as chunked
    # texts, a tag in front of an integer, an indefinite top-level map
    data = (b"\xbf" + T("vectors") + vectors + T("future_field") + b"\x82\x01\xa1\x61a\xf6" +
            T("end_idx") + U(29) + b"\x7f" + T("chunk") + T("_id") + b"\xff" + b"\x7f" + T("chunk") + T("-2") + b"\xff" +
            T("start_idx") + b"\xc1" + U(20) + b"\xff")
    ch = decode_vector_chunk(data)
    assert (ch.chunk_id, ch.start_idx, ch.end_idx) == ("chunk-2", 20, 29)
    assert np.array_equal(ch.ids, ids) and _same_bits(ch.rows, rows)
    o = O.decode_chunk(data)
    assert o[:3] == ("chunk-2", 20, 29) and np.array_equal(o[3], ids) and _same_bits(o[4], rows)


# ---- failures ------------------------------------------------------------------------------------------
def test_malformed_input_is_an_error_not_a_crash():
    ids = _ids(4, 9)
    rows = np.arange(16, dtype=np.float32).reshape(4, 4) + np.float32(0.3)
    good = encode_vector_chunk("c", 1, 2, ids, rows)
    for cut in list(range(0, len(good), 7)) + [len(good) - 1]:
        with pytest.raises(ChunkError):
            decode_vector_chunk(good[:cut])
        with pytest.raises(O.CborError):
            O.decode_chunk(good[:cut])
    with pytest.raises(ChunkError):
        decode_vector_chunk(good + b"\x00")                       # trailing bytes
    with pytest.raises(ChunkError):
        decode_vector_chunk(b"\xff\xff\xff")                      # tests/hnsw/persistence.rs:264's junk
    with pytest.raises(ChunkError):
        decode_vector_chunk(b"\x84\x66\x53\x35")                  # an array, not a VectorChunk
    with pytest.raises(ChunkError):                               # missing field
        decode_vector_chunk(b"\xa1" + O.encode_text("chunk_id") + O.encode_text("x"))
    bad_id = good.replace(b"\x98\x20", b"\x98\x1f", 1)            # 31-byte id: structure no longer parses as a chunk
    with pytest.raises(ChunkError):
        decode_vector_chunk(bad_id)
    rng = np.random.default_rng(0)
    for _ in range(300):                                          # random corruption never crashes
        b = bytearray(good)
        for _ in range(3):
            b[rng.integers(0, len(b))] = rng.integers(0, 256)
        try:
            decode_vector_chunk(bytes(b))
        except (ChunkError, InconsistentDimensions):
            pass


def test_inconsistent_dimensions():
    """IVFError::InconsistentDimensions (src/ivf/core.rs:14-39) — the check `train` / `batch_insert` make per row."""
    a = O.encode_chunk("c", 0, 1, _ids(1, 1), np.ones((1, 4), np.float32))
    extra = b"\x98\x20" + b"".join(O.encode_uint(int(b)) for b in _ids(1, 2)[0]) + b"\x83\xf9\x3c\x00\xf9\x3c\x00\xf9\x3c\x00"
    two = a.replace(O.encode_text("vectors") + b"\xa1", O.encode_text("vectors") + b"\xa2", 1) + extra
    with pytest.raises(InconsistentDimensions):
        decode_vector_chunk(two)
    with pytest.raises(O.CborError):
        O.decode_chunk(two)


def test_capacity_and_argument_errors():
    lib = L.load()
    ids = _ids(3, 4)
    rows = np.ones((3, 8), np.float32) * np.float32(0.7)
    data = encode_vector_chunk("c", 0, 2, ids, rows)
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
    info = L.ChunkInfo()
    assert lib.fvdb_chunk_decode(C.addressof(buf), len(data), C.byref(info), None, None, 0) == L.OK
    assert (info.n_vectors, info.dim, info.chunk_id) == (3, 8, b"c")
    out_ids = np.zeros((2, 32), np.uint8)
    out_rows = np.zeros((2, 8), np.float32)
    rc = lib.fvdb_chunk_decode(C.addressof(buf), len(data), C.byref(info), out_ids.ctypes.data, out_rows.ctypes.data, 2)
    assert rc == L.ERR_INVALID_ARG and b"cap_vectors" in lib.fvdb_chunk_last_error()
    rc = lib.fvdb_chunk_decode(C.addressof(buf), len(data), C.byref(info), out_ids.ctypes.data, None, 2)
    assert rc == L.ERR_INVALID_ARG
    need = C.c_size_t(0)
    small = np.zeros(10, np.uint8)
    rc = lib.fvdb_chunk_encode(b"c", 0, 2, ids.ctypes.data, rows.ctypes.data, 3, 8, small.ctypes.data, small.size, C.byref(need))
    assert rc == L.ERR_INVALID_ARG and need.value == len(data)


# ---- committed fixtures (tests/golden/vector_chunk_*.cbor, made by tests/golden/make_chunk_fixture.py) ----
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_committed_chunk_fixtures(name):
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    data = open(os.path.join(gold, f"vector_chunk_{name}.cbor"), "rb").read()
    z = np.load(os.path.join(gold, "vector_chunk.npz"))
    ids, rows = z[f"{name}_ids"], z[f"{name}_rows"]
    start, end = (int(v) for v in z[f"{name}_meta"])
    cid = bytes(z[f"{name}_chunk_id"]).decode()
    ch = decode_vector_chunk(data)                      # native reader
    assert (ch.chunk_id, ch.start_idx, ch.end_idx) == (cid, start, end)
    assert np.array_equal(ch.ids, ids) and _same_bits(ch.rows, rows)
    o = O.decode_chunk(data)                            # oracle reader
    assert o[:3] == (cid, start, end) and np.array_equal(o[3], ids) and _same_bits(o[4], rows)
    if name != "c":                                     # canonical spelling: both writers reproduce the bytes
        assert encode_vector_chunk(cid, start, end, ids, rows) == data
        assert O.encode_chunk(cid, start, end, ids, rows) == data
