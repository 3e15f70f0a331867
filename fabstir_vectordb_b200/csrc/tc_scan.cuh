// tc_scan.cuh — interface of the tensor-core (tcgen05 / TMEM / TMA) scan path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "common.cuh"

namespace fvdb {

struct TcScratchImpl;

// Per-handle state of the tensor-core path: row norms, TMA descriptors, work buffers.
// The *_dirty flags are raised by the engine whenever the corresponding matrix moved.
struct TcScratch {
    bool arena_dirty = true;
    bool flat_dirty = true;
    bool centroids_dirty = true;
    TcScratchImpl* impl = nullptr;
};

struct TcSearchArgs {
    const float* centroids;
    uint32_t nlist;
    const float* rows;        // IVF arena
    const uint32_t* ids;
    uint64_t n_rows;
    const uint32_t* list_off; // [nlist + 1] device
    const float* Q;
    uint32_t nq, D, k, nprobe;
    const uint64_t* tomb;
    uint64_t tomb_bits;
    const uint64_t* filt;
    uint64_t filt_bits;
    uint32_t shortlist;       // 0 = default
    uint64_t* out_keys;       // [nq][k] exact keys, sorted
    uint64_t* d_scanned_rows; // device counter (distinct posting-list rows streamed)
    cudaEvent_t ev_scan0, ev_scan1;
    int sm_count;
};

struct TcFlatArgs {
    const float* rows;
    const uint32_t* ids;
    uint64_t n_rows;
    const float* Q;
    uint32_t nq, D, k;
    const uint64_t* tomb;
    uint64_t tomb_bits;
    const uint64_t* filt;
    uint64_t filt_bits;
    uint32_t shortlist;
    uint64_t* out_keys;
    int sm_count;
};

// dim % 32 == 0 (one 128-byte swizzle atom per k-block) and dim <= 1024
bool tc_supported(uint32_t D);

int tc_ivf_search(TcScratch& s, const TcSearchArgs& a, cudaStream_t st, size_t* dev_bytes,
                  uint32_t* launches, uint32_t* fallback_queries, std::string* err);
int tc_flat_search(TcScratch& s, const TcFlatArgs& a, cudaStream_t st, size_t* dev_bytes,
                   uint32_t* launches, uint32_t* fallback_queries, std::string* err);
// keys[i] = (exact distance bits << 32) | nearest centroid, exact strict-'<' argmin semantics
int tc_assign(TcScratch& s, const float* centroids, uint32_t nlist, const float* x, uint64_t n,
              uint32_t D, uint64_t* keys, cudaStream_t st, size_t* dev_bytes, std::string* err);
void tc_release(TcScratch& s);

}  // namespace fvdb
