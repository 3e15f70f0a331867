/*
 * fvdb_synth.h — counter-based synthetic data generator (bench / test support; NOT part of the
 * reference-facing ABI).  The reference's own generator is degenerate (1000 distinct vectors,
 * tests/integration/large_dataset_tests.rs:27-38), so benchmarks use a mixture of
 * Gaussians in `dim` dimensions, L2-normalised (MiniLM-style unit vectors), SURVEY §8(d).
 * Every value is a pure function of (seed, row, column) built from integer hashing and
 * correctly-rounded fp32 operations only, so the numpy twin (fabstir_vectordb_b200/synth.py)
 * and every GPU shard produce identical bits without shipping data.
 */
#ifndef FVDB_SYNTH_H_
#define FVDB_SYNTH_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

/* rows [row0, row0+n) of the database into d_out [n x dim] (device pointer, current device).
 *   row r belongs to component r % n_comp; value = centre[comp] + sigma * noise, normalised. */
int fvdb_synth_rows_device(float *d_out, uint64_t row0, uint64_t n, uint32_t dim, uint32_t n_comp,
                           float sigma, uint64_t seed, void *stream);
/* strided variant: output row i is database row row0 + (i / blk) * blk * stride + (i % blk) */
int fvdb_synth_rows_strided_device(float *d_out, uint64_t row0, uint64_t n, uint32_t dim,
                                   uint32_t n_comp, float sigma, uint64_t seed, uint32_t blk,
                                   uint64_t stride, void *stream);
/* queries [q0, q0+n): database row (hash(seed_q, i) % n_total) + qnoise * noise, re-normalised */
int fvdb_synth_queries_device(float *d_out, uint64_t q0, uint64_t n, uint32_t dim, uint64_t n_total,
                              uint32_t n_comp, float sigma, uint64_t seed, float qnoise,
                              uint64_t seed_q, void *stream);
/* filter bitmap: bit(id) = (hash(seed, id) % mod == 0), ids [0, nbits) ; nbits % 64 == 0 */
int fvdb_synth_filter_device(uint64_t *d_bits, uint64_t nbits, uint32_t mod, uint64_t seed,
                             void *stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
