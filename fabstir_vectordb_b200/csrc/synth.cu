// synth.cu — device twin of fabstir_vectordb_b200/synth.py (see include/fvdb_synth.h).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fvdb_synth.h"

namespace {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
__host__ __device__ __forceinline__ uint64_t stream_key(uint64_t seed, uint64_t stream) {
    return mix64(seed + stream * 0x9E3779B97F4A7C15ull);
}
__host__ __device__ __forceinline__ uint64_t hash2(uint64_t key, uint64_t a, uint64_t b) {
    return mix64(key ^ (a * 0xD1B54A32D192ED03ull + b * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull));
}
// Irwin-Hall(4) of 16-bit uniforms, centred and scaled to unit variance: integer sum is exact,
// one correctly rounded multiply.
__device__ __forceinline__ float gauss(uint64_t h) {
    const int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)(h >> 48);
    return __fmul_rn((float)(s - 131070), 2.6429157e-05f);
}

// unnormalised database row element
__device__ __forceinline__ float row_elem(uint64_t kc, uint64_t kx, uint64_t row, uint32_t comp,
                                          uint32_t d, float sigma) {
    const float c = gauss(hash2(kc, comp, d));
    const float e = gauss(hash2(kx, row, d));
    return __fadd_rn(c, __fmul_rn(sigma, e));
}

// sequential fp32 sum of squares over smem (lane 0), broadcast 1/sqrt
__device__ __forceinline__ float inv_norm_seq(const float* v, uint32_t D, int lane) {
    float inv = 0.f;
    if (lane == 0) {
        float ss = 0.f;
        for (uint32_t d = 0; d < D; ++d) ss = __fadd_rn(ss, __fmul_rn(v[d], v[d]));
        inv = __fdiv_rn(1.0f, __fsqrt_rn(ss));
    }
    return __shfl_sync(0xffffffffu, inv, 0);
}

// row index of output i: row0 + (i / blk) * blk * stride + (i % blk)   (blk = stride = 1: dense)
__global__ void synth_rows_kernel(float* __restrict__ out, uint64_t row0, uint64_t n, uint32_t D,
                                  uint32_t n_comp, float sigma, uint64_t seed, uint32_t blk, uint64_t stride) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* v = sm + (size_t)w * D;
    const uint64_t kc = stream_key(seed, 1), kx = stream_key(seed, 2);
    const uint64_t nw = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + w; i < n; i += nw) {
        const uint64_t row = row0 + (i / blk) * blk * stride + (i % blk);
        const uint32_t comp = (uint32_t)(row % n_comp);
        for (uint32_t d = lane; d < D; d += 32) v[d] = row_elem(kc, kx, row, comp, d, sigma);
        __syncwarp();
        const float inv = inv_norm_seq(v, D, lane);
        for (uint32_t d = lane; d < D; d += 32) out[i * D + d] = __fmul_rn(v[d], inv);
        __syncwarp();
    }
}

__global__ void synth_queries_kernel(float* __restrict__ out, uint64_t q0, uint64_t n, uint32_t D,
                                     uint64_t n_total, uint32_t n_comp, float sigma, uint64_t seed,
                                     float qnoise, uint64_t seed_q) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* v = sm + (size_t)w * D;
    const uint64_t kc = stream_key(seed, 1), kx = stream_key(seed, 2);
    const uint64_t kb = stream_key(seed_q, 3), kn = stream_key(seed_q, 4);
    const uint64_t nw = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + w; i < n; i += nw) {
        const uint64_t qi = q0 + i;
        const uint64_t row = hash2(kb, qi, 0) % n_total;
        const uint32_t comp = (uint32_t)(row % n_comp);
        for (uint32_t d = lane; d < D; d += 32) v[d] = row_elem(kc, kx, row, comp, d, sigma);
        __syncwarp();
        float inv = inv_norm_seq(v, D, lane);
        for (uint32_t d = lane; d < D; d += 32) {
            const float x = __fmul_rn(v[d], inv);
            v[d] = __fadd_rn(x, __fmul_rn(qnoise, gauss(hash2(kn, qi, d))));
        }
        __syncwarp();
        inv = inv_norm_seq(v, D, lane);
        for (uint32_t d = lane; d < D; d += 32) out[i * D + d] = __fmul_rn(v[d], inv);
        __syncwarp();
    }
}

__global__ void synth_filter_kernel(uint64_t* __restrict__ bits, uint64_t nwords, uint32_t mod,
                                    uint64_t seed) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    const uint64_t kf = stream_key(seed, 5);
    uint64_t word = 0;
    for (int b = 0; b < 64; ++b) {
        const uint64_t id = w * 64 + b;
        if (hash2(kf, id, 0) % mod == 0) word |= 1ull << b;
    }
    bits[w] = word;
}

}  // namespace

extern "C" {

int fvdb_synth_rows_strided_device(float* d_out, uint64_t row0, uint64_t n, uint32_t dim, uint32_t n_comp,
                                   float sigma, uint64_t seed, uint32_t blk, uint64_t stride, void* stream);

int fvdb_synth_rows_device(float* d_out, uint64_t row0, uint64_t n, uint32_t dim, uint32_t n_comp,
                           float sigma, uint64_t seed, void* stream) {
    return fvdb_synth_rows_strided_device(d_out, row0, n, dim, n_comp, sigma, seed, 1, 1, stream);
}

int fvdb_synth_rows_strided_device(float* d_out, uint64_t row0, uint64_t n, uint32_t dim, uint32_t n_comp,
                                   float sigma, uint64_t seed, uint32_t blk, uint64_t stride, void* stream) {
    if (n == 0) return 0;
    if (dim == 0 || dim > 8192 || n_comp == 0 || blk == 0 || stride == 0) return -12;
    const size_t smem = (size_t)4 * dim * sizeof(float);
    cudaFuncSetAttribute(synth_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    uint64_t blocks = (n + 3) / 4;
    if (blocks > 148ull * 8) blocks = 148ull * 8;
    synth_rows_kernel<<<(uint32_t)blocks, 128, smem, (cudaStream_t)stream>>>(d_out, row0, n, dim, n_comp,
                                                                            sigma, seed, blk, stride);
    return cudaGetLastError() == cudaSuccess ? 0 : -10;
}

int fvdb_synth_queries_device(float* d_out, uint64_t q0, uint64_t n, uint32_t dim, uint64_t n_total,
                              uint32_t n_comp, float sigma, uint64_t seed, float qnoise,
                              uint64_t seed_q, void* stream) {
    if (n == 0) return 0;
    if (dim == 0 || dim > 8192 || n_comp == 0 || n_total == 0) return -12;
    const size_t smem = (size_t)4 * dim * sizeof(float);
    cudaFuncSetAttribute(synth_queries_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    uint64_t blocks = (n + 3) / 4;
    if (blocks > 148ull * 8) blocks = 148ull * 8;
    synth_queries_kernel<<<(uint32_t)blocks, 128, smem, (cudaStream_t)stream>>>(
        d_out, q0, n, dim, n_total, n_comp, sigma, seed, qnoise, seed_q);
    return cudaGetLastError() == cudaSuccess ? 0 : -10;
}

int fvdb_synth_filter_device(uint64_t* d_bits, uint64_t nbits, uint32_t mod, uint64_t seed, void* stream) {
    if (nbits == 0) return 0;
    if (nbits % 64 != 0 || mod == 0) return -12;
    const uint64_t nwords = nbits / 64;
    synth_filter_kernel<<<(uint32_t)((nwords + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_bits, nwords,
                                                                                            mod, seed);
    return cudaGetLastError() == cudaSuccess ? 0 : -10;
}

}  // extern "C"
