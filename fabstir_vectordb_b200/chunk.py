"""VectorChunk CBOR <-> dense arrays, through the C ABI of include/fvdb_chunk.h.

Host mirror of `VectorChunk::{to_cbor, from_cbor}` (src/core/chunk.rs:78-86) for the bulk paths
(`load_index_chunked` / `save_index_chunked`, src/hybrid/persistence.rs:560-660, :188-330): a chunk
becomes (ids [n][32] uint8, rows [n][dim] float32) — the layout `IVFIndex.batch_insert` /
`fvdb_ivf_add` take — in one native call instead of one HashMap entry at a time.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .engine import FvdbError, InconsistentDimensions


class ChunkError(FvdbError):
    """ChunkError::{Serialization, Deserialization} (src/core/chunk.rs:11-27)."""


@dataclass
class VectorChunk:
    """src/core/chunk.rs:37-43, with the HashMap flattened into two arrays in file order."""
    chunk_id: str
    start_idx: int
    end_idx: int
    ids: np.ndarray   # [n][32] uint8 — VectorId bytes
    rows: np.ndarray  # [n][dim] float32
    _pin: object = None  # keeps a page-locked buffer behind `rows` alive

    def __len__(self) -> int:
        return len(self.ids)


def _raise(rc: int, lib):
    msg = (lib.fvdb_chunk_last_error() or b"").decode(errors="replace")
    if rc == L.ERR_INCONSISTENT_DIM:
        raise InconsistentDimensions(msg)
    raise ChunkError(msg or f"chunk codec error {rc}")


def decode_vector_chunk(data: bytes, *, pinned: bool = False) -> VectorChunk:
    """VectorChunk::from_cbor.  pinned=True decodes the rows straight into a page-locked buffer
    (fvdb_host_alloc), so the following upload needs no staging copy."""
    lib = L.load()
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data) if not isinstance(data, (bytearray, memoryview)) else \
        (C.c_ubyte * len(data)).from_buffer(data)
    info = L.ChunkInfo()
    rc = lib.fvdb_chunk_decode(C.addressof(buf), len(data), C.byref(info), None, None, 0)
    if rc != L.OK:
        _raise(rc, lib)
    n, dim = int(info.n_vectors), int(info.dim)
    ids = np.empty((n, 32), dtype=np.uint8)
    pin = None
    if pinned and n * dim:
        from .engine import PinnedArray
        pin = PinnedArray((n, dim), np.float32)
        rows = pin.array
    else:
        rows = np.empty((n, dim), dtype=np.float32)
    if n:
        rc = lib.fvdb_chunk_decode(C.addressof(buf), len(data), C.byref(info), ids.ctypes.data, rows.ctypes.data, n)
        if rc != L.OK:
            _raise(rc, lib)
    return VectorChunk(info.chunk_id.decode("utf-8", errors="replace"), int(info.start_idx), int(info.end_idx), ids, rows, pin)


def encode_vector_chunk(chunk_id: str, start_idx: int, end_idx: int, ids, rows) -> bytes:
    """VectorChunk::to_cbor for entries given as arrays (written in the given order)."""
    lib = L.load()
    ids = np.ascontiguousarray(ids, dtype=np.uint8).reshape(-1, 32)
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n = len(ids)
    rows = rows.reshape(n, -1) if n else rows.reshape(0, 0)
    dim = rows.shape[1] if n else 0
    cid = chunk_id.encode("utf-8")
    need = C.c_size_t(0)
    rc = lib.fvdb_chunk_encode(cid, start_idx, end_idx, ids.ctypes.data, rows.ctypes.data, n, dim, None, 0, C.byref(need))
    if rc != L.OK:
        _raise(rc, lib)
    out = np.empty(need.value, dtype=np.uint8)
    rc = lib.fvdb_chunk_encode(cid, start_idx, end_idx, ids.ctypes.data, rows.ctypes.data, n, dim,
                               out.ctypes.data, out.size, C.byref(need))
    if rc != L.OK:
        _raise(rc, lib)
    return out.tobytes()
