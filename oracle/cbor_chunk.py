"""Oracle for the VectorChunk codec (TEST INFRASTRUCTURE — only tests/ may import this).

A pure-Python restatement of what the reference does with a chunk:
    VectorChunk { chunk_id: String, start_idx: usize, end_idx: usize,
                  vectors: HashMap<VectorId, Vec<f32>> }          src/core/chunk.rs:37-43
    to_cbor = serde_cbor::to_vec(self), from_cbor = serde_cbor::from_slice   src/core/chunk.rs:78-86
    VectorId([u8; 32])                                                        src/core/types.rs:9-10

serde_cbor ("0.11", Cargo.toml:17) is an un-vendored dependency: it is not under /root/reference and
cannot be run here (no cargo).  Its published behaviour is restated: a struct is a definite map of
text keys in declaration order, integers take their shortest form, a `[u8; 32]` is a definite array
of 32 unsigned ints, a `Vec<f32>` a definite array whose elements are written as binary16 (0xf9) when
the value survives f32 -> f16 -> f32, else as binary32 (0xfa); +-inf and NaN are the fixed half
floats f9 7c00 / f9 fc00 / f9 7e00.  No reference test holds a byte-level CBOR fixture (its chunk
tests are round trips, tests/unit/chunk_tests.rs:38-70,340-360), so byte-level PARITY IS UNPINNED;
the item encodings themselves are pinned by the RFC 8949 Appendix A examples (tests/test_chunk_codec.py).
"""
from __future__ import annotations

import struct

import numpy as np


# ---- encoder -------------------------------------------------------------------------------------
def _head(major: int, arg: int) -> bytes:
    m = major << 5
    if arg < 24:
        return bytes([m | arg])
    if arg <= 0xFF:
        return bytes([m | 24, arg])
    if arg <= 0xFFFF:
        return bytes([m | 25]) + struct.pack(">H", arg)
    if arg <= 0xFFFFFFFF:
        return bytes([m | 26]) + struct.pack(">I", arg)
    return bytes([m | 27]) + struct.pack(">Q", arg)


def encode_uint(v: int) -> bytes:
    return _head(0, v)


def encode_text(s: str) -> bytes:
    b = s.encode("utf-8")
    return _head(3, len(b)) + b


def encode_f32(v) -> bytes:
    """serde_cbor 0.11 Serializer::serialize_f32."""
    v = np.float32(v)
    if np.isinf(v):
        return b"\xf9\x7c\x00" if v > 0 else b"\xf9\xfc\x00"
    if np.isnan(v):
        return b"\xf9\x7e\x00"
    with np.errstate(over="ignore"):
        h = np.float16(v)  # round to nearest even, overflow to inf
    if np.float32(h) == v:
        return b"\xf9" + struct.pack(">H", int(h.view(np.uint16)))
    return b"\xfa" + struct.pack(">I", int(v.view(np.uint32)))


def encode_chunk(chunk_id: str, start_idx: int, end_idx: int, ids: np.ndarray, rows: np.ndarray) -> bytes:
    """ids [n][32] uint8, rows [n][dim] float32, entries written in the given order."""
    ids = np.asarray(ids, dtype=np.uint8).reshape(-1, 32)
    rows = np.asarray(rows, dtype=np.float32)
    rows = rows.reshape(len(ids), -1) if len(ids) else rows.reshape(0, 0)
    out = [_head(5, 4), encode_text("chunk_id"), encode_text(chunk_id), encode_text("start_idx"),
           encode_uint(start_idx), encode_text("end_idx"), encode_uint(end_idx), encode_text("vectors"),
           _head(5, len(ids))]
    for i in range(len(ids)):
        out.append(_head(4, 32))
        out.extend(encode_uint(int(b)) for b in ids[i])
        out.append(_head(4, rows.shape[1]))
        out.extend(encode_f32(x) for x in rows[i])
    return b"".join(out)


# ---- decoder (generic RFC 8949 data model -> Python objects) -----------------------------------
class CborError(ValueError):
    pass


class _Break:
    pass


def _decode_item(buf: bytes, pos: int, depth: int = 0):
    if depth > 64:
        raise CborError("nesting too deep")
    if pos >= len(buf):
        raise CborError("truncated")
    b = buf[pos]
    pos += 1
    major, ai = b >> 5, b & 31
    arg = ai
    if 24 <= ai <= 27:
        n = 1 << (ai - 24)
        if pos + n > len(buf):
            raise CborError("truncated")
        arg = int.from_bytes(buf[pos:pos + n], "big")
        pos += n
    elif 28 <= ai <= 30:
        raise CborError("reserved additional information")
    if major == 0:
        if ai == 31:
            raise CborError("indefinite integer")
        return arg, pos
    if major == 1:
        if ai == 31:
            raise CborError("indefinite integer")
        return -1 - arg, pos
    if major in (2, 3):
        if ai == 31:
            parts = []
            while True:
                if pos >= len(buf):
                    raise CborError("truncated")
                if buf[pos] == 0xFF:
                    pos += 1
                    break
                if buf[pos] >> 5 != major or buf[pos] & 31 == 31:
                    raise CborError("bad chunk of an indefinite string")
                part, pos = _decode_item(buf, pos, depth + 1)
                parts.append(part if major == 2 else part.encode("utf-8"))
            data = b"".join(parts)
        else:
            if pos + arg > len(buf):
                raise CborError("truncated")
            data = buf[pos:pos + arg]
            pos += arg
        return (bytes(data) if major == 2 else bytes(data).decode("utf-8")), pos
    if major == 4:
        items = []
        if ai == 31:
            while True:
                if pos >= len(buf):
                    raise CborError("truncated")
                if buf[pos] == 0xFF:
                    pos += 1
                    break
                v, pos = _decode_item(buf, pos, depth + 1)
                items.append(v)
        else:
            for _ in range(arg):
                v, pos = _decode_item(buf, pos, depth + 1)
                items.append(v)
        return items, pos
    if major == 5:
        pairs = []
        if ai == 31:
            while True:
                if pos >= len(buf):
                    raise CborError("truncated")
                if buf[pos] == 0xFF:
                    pos += 1
                    break
                k, pos = _decode_item(buf, pos, depth + 1)
                v, pos = _decode_item(buf, pos, depth + 1)
                pairs.append((k, v))
        else:
            for _ in range(arg):
                k, pos = _decode_item(buf, pos, depth + 1)
                v, pos = _decode_item(buf, pos, depth + 1)
                pairs.append((k, v))
        return pairs, pos  # a list of pairs: keys may be arrays (unhashable)
    if major == 6:
        if ai == 31:
            raise CborError("indefinite tag")
        return _decode_item(buf, pos, depth + 1)  # tags are transparent here
    # major 7
    if ai == 25:
        return np.float32(np.uint16(arg).view(np.float16)), pos
    if ai == 26:
        return np.uint32(arg).view(np.float32), pos
    if ai == 27:
        return np.float32(np.uint64(arg).view(np.float64)), pos
    if ai == 31:
        raise CborError("unexpected break")
    return {20: False, 21: True, 22: None}.get(arg, arg), pos


def decode_item(buf: bytes):
    v, pos = _decode_item(bytes(buf), 0)
    if pos != len(buf):
        raise CborError("trailing bytes")
    return v


def decode_chunk(buf: bytes):
    """-> (chunk_id, start_idx, end_idx, ids [n][32] uint8, rows [n][dim] float32) in file order."""
    top = decode_item(buf)
    if not isinstance(top, list) or (top and not isinstance(top[0], tuple)):
        raise CborError("a VectorChunk is a map")
    fields = {}
    for k, v in top:
        if not isinstance(k, str):
            raise CborError("field names are text")
        if k in ("chunk_id", "start_idx", "end_idx", "vectors"):
            if k in fields:
                raise CborError("duplicate field " + k)
            fields[k] = v
    if set(fields) != {"chunk_id", "start_idx", "end_idx", "vectors"}:
        raise CborError("missing field")
    if not isinstance(fields["chunk_id"], str):
        raise CborError("chunk_id is text")
    for k in ("start_idx", "end_idx"):
        if not isinstance(fields[k], int) or isinstance(fields[k], bool) or fields[k] < 0:
            raise CborError(k + " is an unsigned integer")
    ids, rows = [], []
    dim = None
    for k, v in fields["vectors"]:
        if isinstance(k, bytes):
            if len(k) != 32:
                raise CborError("a VectorId is 32 bytes")
            ids.append(np.frombuffer(k, dtype=np.uint8))
        else:
            if not isinstance(k, list) or len(k) != 32 or any(not isinstance(b, int) or not 0 <= b <= 255 for b in k):
                raise CborError("a VectorId is 32 bytes")
            ids.append(np.array(k, dtype=np.uint8))
        if not isinstance(v, list):
            raise CborError("a vector is an array")
        if dim is None:
            dim = len(v)
        elif len(v) != dim:
            raise CborError("inconsistent dimensions")
        rows.append(np.array([np.float32(x) for x in v], dtype=np.float32))
    n = len(ids)
    ids_a = np.stack(ids) if n else np.zeros((0, 32), np.uint8)
    rows_a = np.stack(rows).reshape(n, dim or 0) if n else np.zeros((0, 0), np.float32)
    return fields["chunk_id"], fields["start_idx"], fields["end_idx"], ids_a, rows_a
