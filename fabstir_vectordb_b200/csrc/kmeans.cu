// kmeans.cu — k-means pieces of IVFIndex::train (src/ivf/core.rs:240-429) that are not the
// assignment scan (that one is exact_scan.cu / tc kernels).
#include "common.cuh"
#include "kernels.cuh"

namespace fvdb {

namespace {

// update_centroids, src/ivf/core.rs:388-417.  Thread (cluster c, dim d) adds the members of c
// IN DATA ORDER (perm is a stable grouping), so every per-cluster f32 sum is bit-identical to
// the reference's `*s += v` loop; mean = sum / count as f32; empty cluster keeps its centroid.
__global__ void centroid_update_kernel(const float* __restrict__ data, uint32_t D,
                                       const uint32_t* __restrict__ offsets,
                                       const uint32_t* __restrict__ perm,
                                       float* __restrict__ centroids) {
    const uint32_t c = blockIdx.x;
    const uint32_t d = blockIdx.y * blockDim.x + threadIdx.x;
    const uint32_t b = offsets[c], e = offsets[c + 1];
    if (d >= D || e == b) return;
    float sum = 0.0f;
    uint32_t m = b;
    for (; m + 4 <= e; m += 4) {
        const uint32_t i0 = perm[m], i1 = perm[m + 1], i2 = perm[m + 2], i3 = perm[m + 3];
        const float v0 = __ldg(data + (size_t)i0 * D + d);
        const float v1 = __ldg(data + (size_t)i1 * D + d);
        const float v2 = __ldg(data + (size_t)i2 * D + d);
        const float v3 = __ldg(data + (size_t)i3 * D + d);
        sum = __fadd_rn(sum, v0);
        sum = __fadd_rn(sum, v1);
        sum = __fadd_rn(sum, v2);
        sum = __fadd_rn(sum, v3);
    }
    for (; m < e; ++m) sum = __fadd_rn(sum, __ldg(data + (size_t)perm[m] * D + d));
    centroids[(size_t)c * D + d] = __fdiv_rn(sum, (float)(e - b));
}

// dist[i] = euclidean_distance_scalar(data[i], centroids[assign[i]]) in reference order.
// One warp handles 32 points; 32x32 tiles of points and of their centroids are staged through
// shared memory so global reads are coalesced while each lane still walks d sequentially.
__global__ void __launch_bounds__(128) rowwise_dist_kernel(const float* __restrict__ data, uint64_t n,
                                                           uint32_t D, const float* __restrict__ centroids,
                                                           const uint32_t* __restrict__ assign,
                                                           float* __restrict__ dist) {
    __shared__ float xs[4][32][33];
    __shared__ float cs[4][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t base = ((uint64_t)blockIdx.x * 4 + w) * 32;
    if (base >= n) return;
    const uint64_t i = base + lane;
    const uint32_t my_c = (i < n) ? assign[i] : 0;
    float acc = 0.0f;
    for (uint32_t kc = 0; kc < D; kc += 32) {
        // row r of the tile is loaded by the whole warp (lane = column)
        for (int r = 0; r < 32; ++r) {
            const uint64_t p = base + r;
            const uint32_t col = kc + lane;
            const uint32_t cr = __shfl_sync(0xffffffffu, my_c, r);
            float xv = 0.f, cv = 0.f;
            if (p < n && col < D) {
                xv = __ldg(data + (size_t)p * D + col);
                cv = __ldg(centroids + (size_t)cr * D + col);
            }
            xs[w][r][lane] = xv;
            cs[w][r][lane] = cv;
        }
        __syncwarp();
#pragma unroll 8
        for (int d = 0; d < 32; ++d) {
            const float t = __fsub_rn(xs[w][lane][d], cs[w][lane][d]);
            acc = __fadd_rn(acc, __fmul_rn(t, t));
        }
        __syncwarp();
    }
    if (i < n) dist[i] = __fsqrt_rn(acc);
}

__global__ void accumulate_sums_kernel(const float* __restrict__ data, uint64_t n, uint32_t D,
                                       const uint32_t* __restrict__ assign,
                                       const float* __restrict__ dist, float* __restrict__ sums,
                                       uint32_t* __restrict__ counts, double* __restrict__ sqerr) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    double err = 0.0;
    for (uint64_t i = warp; i < n; i += nwarps) {
        const uint32_t a = assign[i];
        const float* x = data + (size_t)i * D;
        float* s = sums + (size_t)a * D;
        for (uint32_t d = lane; d < D; d += 32) atomicAdd(s + d, __ldg(x + d));
        if (lane == 0) {
            atomicAdd(counts + a, 1u);
            if (dist) { const double dd = (double)dist[i]; err += dd * dd; }
        }
    }
    if (lane == 0 && sqerr && err != 0.0) atomicAdd(sqerr, err);
}

__global__ void apply_means_kernel(const float* __restrict__ sums, const uint32_t* __restrict__ counts,
                                   uint32_t nlist, uint32_t D, float* __restrict__ centroids) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)nlist * D) return;
    const uint32_t c = (uint32_t)(i / D);
    const uint32_t cnt = counts[c];
    if (cnt > 0) centroids[i] = __fdiv_rn(sums[i], (float)cnt);
}

__global__ void copy_row_kernel(const float* __restrict__ data, const uint32_t* __restrict__ idx,
                                uint32_t D, float* __restrict__ dst) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D) dst[d] = data[(size_t)(*idx) * D + d];
}

}  // namespace

cudaError_t launch_centroid_update(const float* data, uint32_t D, const uint32_t* offsets,
                                   const uint32_t* perm, uint32_t nlist, float* centroids,
                                   cudaStream_t stream) {
    if (nlist == 0) return cudaSuccess;
    dim3 grid(nlist, (D + 127) / 128);
    centroid_update_kernel<<<grid, 128, 0, stream>>>(data, D, offsets, perm, centroids);
    return cudaGetLastError();
}

cudaError_t launch_rowwise_dist(const float* data, uint64_t n, uint32_t D, const float* centroids,
                                const uint32_t* assign, float* dist, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const uint32_t blocks = (uint32_t)((n + 127) / 128);
    rowwise_dist_kernel<<<blocks, 128, 0, stream>>>(data, n, D, centroids, assign, dist);
    return cudaGetLastError();
}

cudaError_t launch_accumulate_sums(const float* data, uint64_t n, uint32_t D, const uint32_t* assign,
                                   const float* dist, float* sums, uint32_t* counts, double* sqerr,
                                   cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    uint64_t blocks = (n * 32 + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    accumulate_sums_kernel<<<(uint32_t)blocks, 256, 0, stream>>>(data, n, D, assign, dist, sums, counts,
                                                                sqerr);
    return cudaGetLastError();
}

cudaError_t launch_apply_means(const float* sums, const uint32_t* counts, uint32_t nlist, uint32_t D,
                               float* centroids, cudaStream_t stream) {
    const uint64_t n = (uint64_t)nlist * D;
    if (n == 0) return cudaSuccess;
    apply_means_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, stream>>>(sums, counts, nlist, D, centroids);
    return cudaGetLastError();
}

cudaError_t launch_copy_row(const float* data, const uint32_t* idx, uint32_t D, float* dst,
                            cudaStream_t stream) {
    copy_row_kernel<<<(D + 127) / 128, 128, 0, stream>>>(data, idx, D, dst);
    return cudaGetLastError();
}

}  // namespace fvdb
