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

// Metrics other than L2 rank by a SIMILARITY, best = largest (dot_product_scalar / cosine_similarity_scalar,
// src/core/vector_ops.rs:35-49, ranked by top_k_indices :12-23: descending, ties in index order).  Their
// candidates use the same u64 keys, ascending = best first: the upper word is an order-preserving image of
// -similarity (the usual sign-flip of the f32 bits), so every selection, merge and tie rule of the L2 path
// applies unchanged; -0.0 is folded onto +0.0 first (the reference compares them equal).
constexpr int METRIC_L2 = 0, METRIC_COS = 1, METRIC_DOT = 2;
__host__ __device__ __forceinline__ uint32_t sim_to_key32(float s) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(-(s + 0.0f));
#else
    union { float f; uint32_t u; } c;
    c.f = -(s + 0.0f);
    uint32_t b = c.u;
#endif
    if (b == 0x80000000u) b = 0u;   // -(+0) = -0: fold onto +0
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ float key32_to_sim(uint32_t k) {
    const uint32_t b = k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu);
    // 0 - x, not -x: the image of a zero similarity decodes to +0.0, as the scalar kernels return it
#ifdef __CUDA_ARCH__
    return 0.0f - __uint_as_float(b);
#else
    union { float f; uint32_t u; } c;
    c.u = b;
    return 0.0f - c.f;
#endif
}
// the value a caller sees for a key: the true L2 distance, or the similarity
__device__ __forceinline__ float key_value(uint64_t k, int metric) {
    return metric == METRIC_L2 ? __uint_as_float((uint32_t)(k >> 32)) : key32_to_sim((uint32_t)(k >> 32));
}

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
    uint32_t sub;         // row-range index when a long posting list is split into several items (else 0)
};

struct DeviceError {
    int nan_flag;       // set when a NaN was seen in an input
    int overflow_flag;  // internal capacity exceeded (bug guard)
};

#ifdef __CUDACC__
// Exact L2 in the reference's operation order for one row per lane: this lane walks its own row
// (16-byte loads, eight in flight) against the query staged in shared memory; the accumulation is
// the strictly sequential f32 chain of euclidean_distance_scalar (src/core/vector_ops.rs:51-57).
__device__ __forceinline__ float exact_l2_lane(const float* __restrict__ q_s, const float* __restrict__ row, uint32_t D) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const float4* q4 = reinterpret_cast<const float4*>(q_s);
    float acc = 0.0f;
    const uint32_t n4 = D >> 2;
    uint32_t i = 0;
    for (; i + 8 <= n4; i += 8) {
        float4 x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __ldg(r4 + i + j);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 qq = q4[i + j];
            float t;
            t = __fsub_rn(qq.x, x[j].x); acc = __fadd_rn(acc, __fmul_rn(t, t));
            t = __fsub_rn(qq.y, x[j].y); acc = __fadd_rn(acc, __fmul_rn(t, t));
            t = __fsub_rn(qq.z, x[j].z); acc = __fadd_rn(acc, __fmul_rn(t, t));
            t = __fsub_rn(qq.w, x[j].w); acc = __fadd_rn(acc, __fmul_rn(t, t));
        }
    }
    for (; i < n4; ++i) {
        const float4 xx = __ldg(r4 + i), qq = q4[i];
        float t;
        t = __fsub_rn(qq.x, xx.x); acc = __fadd_rn(acc, __fmul_rn(t, t));
        t = __fsub_rn(qq.y, xx.y); acc = __fadd_rn(acc, __fmul_rn(t, t));
        t = __fsub_rn(qq.z, xx.z); acc = __fadd_rn(acc, __fmul_rn(t, t));
        t = __fsub_rn(qq.w, xx.w); acc = __fadd_rn(acc, __fmul_rn(t, t));
    }
    return __fsqrt_rn(acc);
}


// two rows per lane in lockstep (two independent accumulation chains: same bits, twice the ILP)
__device__ __forceinline__ void exact_l2_lane2(const float* __restrict__ q_s, const float* __restrict__ row0,
                                               const float* __restrict__ row1, uint32_t D, float& d0, float& d1) {
    const float4* a4 = reinterpret_cast<const float4*>(row0);
    const float4* b4 = reinterpret_cast<const float4*>(row1);
    const float4* q4 = reinterpret_cast<const float4*>(q_s);
    float acc0 = 0.0f, acc1 = 0.0f;
    const uint32_t n4 = D >> 2;
    uint32_t i = 0;
    for (; i + 8 <= n4; i += 8) {
        float4 x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { x[j] = __ldg(a4 + i + j); y[j] = __ldg(b4 + i + j); }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 qq = q4[i + j];
            float t, u;
            t = __fsub_rn(qq.x, x[j].x); u = __fsub_rn(qq.x, y[j].x);
            acc0 = __fadd_rn(acc0, __fmul_rn(t, t)); acc1 = __fadd_rn(acc1, __fmul_rn(u, u));
            t = __fsub_rn(qq.y, x[j].y); u = __fsub_rn(qq.y, y[j].y);
            acc0 = __fadd_rn(acc0, __fmul_rn(t, t)); acc1 = __fadd_rn(acc1, __fmul_rn(u, u));
            t = __fsub_rn(qq.z, x[j].z); u = __fsub_rn(qq.z, y[j].z);
            acc0 = __fadd_rn(acc0, __fmul_rn(t, t)); acc1 = __fadd_rn(acc1, __fmul_rn(u, u));
            t = __fsub_rn(qq.w, x[j].w); u = __fsub_rn(qq.w, y[j].w);
            acc0 = __fadd_rn(acc0, __fmul_rn(t, t)); acc1 = __fadd_rn(acc1, __fmul_rn(u, u));
        }
    }
    for (; i < n4; ++i) {
        const float4 xx = __ldg(a4 + i), yy = __ldg(b4 + i), qq = q4[i];
        float t, u;
        t = __fsub_rn(qq.x, xx.x); u = __fsub_rn(qq.x, yy.x);
        acc0 = __fadd_rn(acc0, __fmul_rn(t, t)); acc1 = __fadd_rn(acc1, __fmul_rn(u, u));
        t = __fsub_rn(qq.y, xx.y); u = __fsub_rn(qq.y, yy.y);
        acc0 = __fadd_rn(acc0, __fmul_rn(t, t)); acc1 = __fadd_rn(acc1, __fmul_rn(u, u));
        t = __fsub_rn(qq.z, xx.z); u = __fsub_rn(qq.z, yy.z);
        acc0 = __fadd_rn(acc0, __fmul_rn(t, t)); acc1 = __fadd_rn(acc1, __fmul_rn(u, u));
        t = __fsub_rn(qq.w, xx.w); u = __fsub_rn(qq.w, yy.w);
        acc0 = __fadd_rn(acc0, __fmul_rn(t, t)); acc1 = __fadd_rn(acc1, __fmul_rn(u, u));
    }
    d0 = __fsqrt_rn(acc0);
    d1 = __fsqrt_rn(acc1);
}
#endif  // __CUDACC__

inline int div_up(int a, int b) { return (a + b - 1) / b; }
inline uint64_t div_up64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

}  // namespace fvdb
