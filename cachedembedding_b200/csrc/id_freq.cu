// id -> frequency histogram over a stream of KJT id batches: the GPU counterpart of the reference's np.bincount
// counters (recsys/datasets/feature_counter.py:21-29,41-60), whose output (int64[N], one count per global id) feeds
// DATASET eviction and the LFU warm start (CachedParamMgr.reorder).
// Integer work, HBM-atomic bound.  Ids arrive feature-major, so the 32 ids of a warp belong to one table and, for the
// small tables, are mostly the same few rows: one lane per distinct id adds the warp's multiplicity.
#include "common.cuh"
#include "profile.cuh"

namespace cebag {
namespace {

__global__ void __launch_bounds__(256)
id_histogram_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t* __restrict__ freq, int64_t num_rows,
                    int32_t* __restrict__ bad) {
    const int lane = lane_id();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (n + stride - 1) / stride;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = 0; r < rounds; ++r, i += stride) {           // whole warps stay in the loop (match.any below)
        long long id = i < n ? ids[i] : -1;
        const bool ok = id >= 0 && id < num_rows;
        if (i < n && !ok) *bad = 1;
        const unsigned peers = __match_any_sync(0xffffffffu, ok ? id : -1 - (long long)lane);
        if (ok && lane == __ffs(peers) - 1)
            atomicAdd(reinterpret_cast<unsigned long long*>(freq + id), (unsigned long long)__popc(peers));
    }
}

}  // namespace
}  // namespace cebag

using namespace cebag;

extern "C" int cebag_id_histogram(const int64_t* ids, int64_t n, int64_t* freq, int64_t num_rows, int32_t* bad_flag,
                                  void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(n >= 0 && num_rows > 0, "sizes");
    if (n == 0) return CEBAG_OK;
    CEBAG_REQUIRE(ids != nullptr && freq != nullptr && bad_flag != nullptr, "ids / freq / bad_flag");
    KernelScope scope(kKernIdHistogram, stream);
    id_histogram_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(ids, n, freq, num_rows, bad_flag);
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}
