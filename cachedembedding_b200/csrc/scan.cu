// Device-wide exclusive scan of int32 arrays, used by the miss-bitmap ranking, the free-slot compaction and the
// radix sort's bucket offsets.  Integer-only, HBM/L2-bound; arrays here are small (<= a few M entries).
#include "common.cuh"
#include "scan.cuh"
#include "profile.cuh"

namespace cebag {

namespace {

constexpr int kChunkItems = 16;                               // ints per thread in the chunked kernels
constexpr int kChunk = kScanThreads * kChunkItems;            // 4096 entries per CTA

// S2 / small arrays: one CTA walks the array with a running carry.
__global__ void __launch_bounds__(1024) scan_single_cta_kernel(int32_t* data, int64_t len, int32_t* total_out) {
    __shared__ int32_t warp_sums[32];
    __shared__ int32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < len; base += 1024) {
        int64_t i = base + threadIdx.x;
        int32_t v = i < len ? data[i] : 0;
        int32_t excl = block_exclusive_scan<1024>(v, warp_sums);
        int32_t carry = carry_s;
        if (i < len) data[i] = carry + excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

// S1: per-chunk totals
__global__ void __launch_bounds__(kScanThreads) chunk_totals_kernel(const int32_t* __restrict__ data, int64_t len,
                                                                  int32_t* __restrict__ totals) {
    __shared__ int32_t warp_sums[32];
    int64_t base = (int64_t)blockIdx.x * kChunk;
    int32_t s = 0;
#pragma unroll
    for (int k = 0; k < kChunkItems; ++k) {
        int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
        if (i < len) s += data[i];
    }
    int32_t tot = block_reduce_sum<kScanThreads>(s, warp_sums);
    if (threadIdx.x == 0) totals[blockIdx.x] = tot;
}

// S3: local exclusive scan of a chunk (blocked arrangement) + chunk base
__global__ void __launch_bounds__(kScanThreads) chunk_downsweep_kernel(int32_t* __restrict__ data, int64_t len,
                                                                     const int32_t* __restrict__ chunk_base) {
    __shared__ int32_t warp_sums[32];
    int64_t base = (int64_t)blockIdx.x * kChunk + (int64_t)threadIdx.x * kChunkItems;
    int32_t v[kChunkItems];
    int32_t s = 0;
#pragma unroll
    for (int k = 0; k < kChunkItems; ++k) {
        v[k] = (base + k < len) ? data[base + k] : 0;
        s += v[k];
    }
    int32_t excl = block_exclusive_scan<kScanThreads>(s, warp_sums) + chunk_base[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kChunkItems; ++k) {
        if (base + k < len) data[base + k] = excl;
        excl += v[k];
    }
}

}  // namespace

size_t scan_workspace_bytes(int64_t len) {
    return (size_t)(ceil_div(len, kChunk) + 1) * sizeof(int32_t);
}

int exclusive_scan_inplace(int32_t* data, int64_t len, int32_t* total_out, int32_t* workspace, cudaStream_t stream) {
    if (len <= 0) {
        if (total_out) CEBAG_CUDA_CHECK(cudaMemsetAsync(total_out, 0, sizeof(int32_t), stream));
        return CEBAG_OK;
    }
    if (len <= 16384) {
        scan_single_cta_kernel<<<1, 1024, 0, stream>>>(data, len, total_out);
        count_launches(1);
        CEBAG_LAUNCH_CHECK();
        return CEBAG_OK;
    }
    int64_t chunks = ceil_div(len, kChunk);
    CEBAG_REQUIRE(chunks <= (int64_t)1 << 22, "scan too long");
    chunk_totals_kernel<<<(int)chunks, kScanThreads, 0, stream>>>(data, len, workspace);
    CEBAG_LAUNCH_CHECK();
    scan_single_cta_kernel<<<1, 1024, 0, stream>>>(workspace, chunks, total_out);
    CEBAG_LAUNCH_CHECK();
    chunk_downsweep_kernel<<<(int)chunks, kScanThreads, 0, stream>>>(data, len, workspace);
    count_launches(3);
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}

}  // namespace cebag
