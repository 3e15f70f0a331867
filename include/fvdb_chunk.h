/* fvdb_chunk.h — VectorChunk CBOR <-> the staging layout the engine uploads.
 *
 * Replaces, for the bulk-load / bulk-save path, the per-row work of
 *   VectorChunk::from_cbor / to_cbor          src/core/chunk.rs:78-86
 *   HybridPersister::load_index_chunked       src/hybrid/persistence.rs:560-660  (decode, then one
 *       find_cluster per vector inside a loop over clusters, :626-653)
 *   HybridPersister::save_index_chunked       src/hybrid/persistence.rs:188-330  (10 K vectors per chunk, :189)
 * A chunk is decoded ONCE into two dense arrays — the 32-byte VectorIds in file order and the
 * vectors as row-major fp32 — which are exactly what fvdb_assign / fvdb_ivf_add / fvdb_flat_add take
 * (the buffers may come from fvdb_host_alloc, so the upload needs no staging copy).
 *
 * Wire format (serde_cbor 0.11, `#[derive(Serialize)]` on `VectorChunk`, src/core/chunk.rs:37-43):
 *   map(4) { "chunk_id": text, "start_idx": uint, "end_idx": uint,
 *            "vectors": map(n) { array(32) of uint (VectorId = [u8; 32], src/core/types.rs:9-10)
 *                                  -> array(dim) of float } }
 * serde_cbor writes an f32 as a half float (0xf9) when that is lossless, else as 0xfa; the
 * decoder accepts f16 / f32 / f64 and integers, definite and indefinite lengths, keys in any order,
 * unknown keys (skipped) and a 32-byte byte string in place of the id array.  HashMap iteration
 * order is unspecified in the reference, so "file order" carries no meaning beyond this call.
 *
 * Pure host functions: no handle, no CUDA call, thread-safe.  Return 0 or a negative fvdb_status
 * (include/fvdb.h): FVDB_ERR_CHUNK_LOAD for malformed / truncated input (ChunkError::Deserialization),
 * FVDB_ERR_INCONSISTENT_DIM when the vectors of one chunk differ in length, FVDB_ERR_INVALID_ARG
 * when an output buffer is too small.
 */
#ifndef FVDB_CHUNK_H
#define FVDB_CHUNK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */

typedef struct fvdb_chunk_info {
    char chunk_id[128];  /* NUL-terminated; longer ids are truncated */
    uint64_t start_idx;  /* VectorChunk::start_idx */
    uint64_t end_idx;    /* VectorChunk::end_idx */
    uint64_t n_vectors;  /* entries of `vectors` */
    uint32_t dim;        /* length of every vector; 0 for an empty chunk */
} fvdb_chunk_info;

/* Decode one chunk.  out_ids [cap_vectors][32] and out_rows [cap_vectors][dim] may both be NULL: only
 * *info is filled (sizing pass).  Otherwise cap_vectors >= n_vectors is required. */
int fvdb_chunk_decode(const uint8_t *cbor, size_t len, fvdb_chunk_info *info, uint8_t *out_ids,
                      float *out_rows, uint64_t cap_vectors);

/* Encode n vectors as one chunk, byte for byte what serde_cbor 0.11 emits for the same entries in
 * the same order.  out may be NULL: only *out_len (the exact size) is computed. */
int fvdb_chunk_encode(const char *chunk_id, uint64_t start_idx, uint64_t end_idx, const uint8_t *ids,
                      const float *rows, uint64_t n, uint32_t dim, uint8_t *out, size_t cap,
                      size_t *out_len);

/* Message of the last failed fvdb_chunk_* call of the calling thread ("" if none). */
const char *fvdb_chunk_last_error(void);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* FVDB_CHUNK_H */
