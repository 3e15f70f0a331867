// common.cuh — shared device/host helpers of libfvdb_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace fvdb {

// A search candidate is one u64 key: (f32 bits of the true L2 distance) << 32 | row id.
// L2 distances are >= +0 so the bit pattern is monotone and u64 order == (distance, id)
// order — the canonical tie rule of include/fvdb.h.  KEY_NONE marks "no candidate".
constexpr uint64_t KEY_NONE = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t ID_NONE = 0xFFFFFFFFu;

__host__ __device__ __forceinline__ uint64_t make_key(float dist, uint32_t id) {
#ifdef __CUDA_ARCH__
    return ((uint64_t)__float_as_uint(dist) << 32) | id;
#else
    union { float f; uint32_t u; } c;
    c.f = dist;
    return ((uint64_t)c.u << 32) | id;
#endif
}
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) { return (uint32_t)k; }
__device__ __forceinline__ float key_dist(uint64_t k) { return __uint_as_float((uint32_t)(k >> 32)); }

// bitmap over row ids (u64 words).  nbits bounds the bitmap; ids beyond it read as 0.
__device__ __forceinline__ bool bit_test(const uint64_t* __restrict__ bits, uint64_t nbits,
                                         uint32_t id) {
    if (id >= nbits) return false;
    return (__ldg(bits + (id >> 6)) >> (id & 63)) & 1ull;
}

// One work item of the exact scan: rows [row_begin,row_end) of a row matrix against a group
// of queries, producing one sorted partial top-k per query.
struct ScanItem {
    uint32_t row_begin;
    uint32_t row_end;
    uint32_t pair_begin;  // first (query,slot) pair, or first query index when identity
    uint32_t pair_count;  // queries in this item (<= TQ)
    uint32_t slot;        // partial slot when identity
    uint32_t identity;    // 1: query = pair_begin + i, slot = slot; 0: via pair arrays
};

struct DeviceError {
    int nan_flag;       // set when a NaN was seen in an input
    int overflow_flag;  // internal capacity exceeded (bug guard)
};

inline int div_up(int a, int b) { return (a + b - 1) / b; }
inline uint64_t div_up64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

}  // namespace fvdb
