// Block-level scan/reduce helpers + the device-wide exclusive scan entry point (scan.cu).
#pragma once
#include "common.cuh"

namespace cebag {

constexpr int kScanThreads = 256;

__device__ __forceinline__ int32_t warp_inclusive_scan(int32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}

// exclusive prefix of `v` over the THREADS threads of the CTA (thread order); warp_sums: >= THREADS/32 ints of smem.
// Ends with a __syncthreads-protected state: callers may reuse warp_sums after their next __syncthreads.
template <int THREADS>
__device__ __forceinline__ int32_t block_exclusive_scan(int32_t v, int32_t* warp_sums) {
    constexpr int W = THREADS / 32;
    int32_t incl = warp_inclusive_scan(v);
    int warp = threadIdx.x >> 5;
    __syncthreads();  // previous users of warp_sums are done
    if (lane_id() == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int32_t w = lane_id() < W ? warp_sums[lane_id()] : 0;
        int32_t wi = warp_inclusive_scan(w);
        if (lane_id() < W) warp_sums[lane_id()] = wi - w;  // exclusive warp base
    }
    __syncthreads();
    return warp_sums[warp] + incl - v;
}

template <int THREADS>
__device__ __forceinline__ int32_t block_reduce_sum(int32_t v, int32_t* warp_sums) {
    constexpr int W = THREADS / 32;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if (lane_id() == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    int32_t t = 0;
    if (threadIdx.x < 32) {
        t = lane_id() < W ? warp_sums[lane_id()] : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    }
    return t;  // valid in warp 0 (all its lanes)
}

// In-place exclusive scan of data[len]; grand total to *total_out (device pointer, may be null).
// workspace: scan_workspace_bytes(len) bytes of device memory.
size_t scan_workspace_bytes(int64_t len);
int exclusive_scan_inplace(int32_t* data, int64_t len, int32_t* total_out, int32_t* workspace, cudaStream_t stream);

}  // namespace cebag
