// Forward of the embedding bag as a TMA-driven row gather (bulk asynchronous copies, sm_90+/sm_100a):
//   cache rows --cp.async.bulk (global -> shared, mbarrier complete_tx)--> a shared-memory stage of kRows rows
//   stage      --cp.async.bulk (shared -> global, bulk group)-----------> kRows CONSECUTIVE output rows (one copy)
// The reference's DLRM call has pooling factor 1 (one id per feature per sample, recsys/datasets/criteo.py:129-130,
// mode sum, no weights), i.e. out[g] = cache[slot[g]]: the kernel is then a pure gather -- no lane ever touches the row
// data, one thread issues a 512 B row load and one thread per tile issues the 16 KB store, so the issue slots that
// the LDG.128 / STG.128 formulation spends on moving bytes (ncu: 42 % issue-slot-busy, 76 M warp instructions per
// launch) are free and the SM keeps ~100 KB of loads and stores in flight.  Tiles that contain a bag with more or
// fewer than one entry (or an invalid slot) are summed by the warp into the same stage and leave through the same
// bulk store, so the kernel is correct for any offsets in mode sum without per-sample weights.
// Each warp owns kStages stages and its own mbarriers: warps never synchronise with each other.
#include "bag_common.cuh"
#include "profile.cuh"

namespace cebag {

namespace {

constexpr int kTmaWarps = 2;                 // warps per CTA
constexpr int kTmaRows = 32;                 // rows (bags) per stage: lane l owns bag tile * 32 + l
constexpr int kTmaStages = 3;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src_smem), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}

// row_bytes: bytes of one row (multiple of 16, <= 2048 so that a stage of 32 rows fits 64 KB)
__global__ void __launch_bounds__(kTmaWarps * 32)
bag_forward_tma_kernel(const BagParams p, float* __restrict__ out, int row_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[kTmaWarps][kTmaStages];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int row_f4 = row_bytes / 16;
    const size_t stage_bytes = (size_t)kTmaRows * row_bytes;
    unsigned char* my_smem = smem + (size_t)warp * kTmaStages * stage_bytes;
    if (lane == 0) {
        for (int s = 0; s < kTmaStages; ++s) mbar_init(smem_addr(&bars[warp][s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t num_tiles = (p.num_bags + kTmaRows - 1) / kTmaRows;
    const int64_t gwarp = (int64_t)blockIdx.x * kTmaWarps + warp;
    const int64_t nwarps = (int64_t)gridDim.x * kTmaWarps;
    const char* cache = reinterpret_cast<const char*>(p.cache);

    // software pipeline over this warp's tiles: `issue` runs kStages - 1 tiles ahead of `drain`
    int64_t issue_tile = gwarp, drain_tile = gwarp;
    int issue_slot = 0, drain_slot = 0;
    uint32_t parity_bits = 0;                      // phase parity of every stage's mbarrier
    uint32_t plain_bits = 0;                       // stages whose tile was summed by the warp (no bulk loads pending)
    int in_flight = 0;

    auto issue = [&](int64_t tile, int slot) {
        const int64_t g = tile * kTmaRows + lane;
        const bool have = g < p.num_bags;
        int64_t lo = 0, hi = 0;
        if (have) { lo = load_offset(p, g); hi = load_offset(p, g + 1); }
        long long slot_id = -1;
        if (hi - lo == 1) slot_id = __ldg(p.slot_ids + lo);
        const bool simple = !have || (hi - lo == 1 && slot_id >= 0 && slot_id < p.cache_rows && slot_id != p.padding_idx);
        const bool all_simple = __all_sync(0xffffffffu, simple);
        const int rows = (int)min((int64_t)kTmaRows, p.num_bags - tile * kTmaRows);
        unsigned char* stage = my_smem + (size_t)slot * stage_bytes;
        // the bulk store that last read this stage must be done with it
        if (lane == 0) bulk_wait_read<1>();     // exactly one younger store (the previous tile's) may still be reading
        __syncwarp();
        if (all_simple) {
            const uint32_t bar = smem_addr(&bars[warp][slot]);
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)rows * row_bytes);
            __syncwarp();
            if (have) bulk_load(smem_addr(stage + (size_t)lane * row_bytes), cache + (size_t)slot_id * row_bytes, row_bytes, bar);
            plain_bits &= ~(1u << slot);
        } else {
            // general bags: the warp sums every bag of the tile into the stage (128-bit columns strided over the lanes)
            for (int r = 0; r < rows; ++r) {
                const int64_t rlo = __shfl_sync(0xffffffffu, lo, r), rhi = __shfl_sync(0xffffffffu, hi, r);
                for (int c = lane; c < row_f4; c += 32) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int64_t i = rlo; i < rhi; ++i) {
                        const long long s = __ldg(p.slot_ids + i);
                        if (s < 0 || s >= p.cache_rows || s == p.padding_idx) continue;
                        add4(acc, ld_stream_f4(reinterpret_cast<const float4*>(cache + (size_t)s * row_bytes) + c));
                    }
                    reinterpret_cast<float4*>(stage + (size_t)r * row_bytes)[c] = acc;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the bulk store
            plain_bits |= 1u << slot;
        }
        __syncwarp();
    };
    auto drain = [&](int64_t tile, int slot) {
        const int rows = (int)min((int64_t)kTmaRows, p.num_bags - tile * kTmaRows);
        if (!((plain_bits >> slot) & 1u)) {
            if (lane == 0) mbar_wait(smem_addr(&bars[warp][slot]), (parity_bits >> slot) & 1u);
            parity_bits ^= 1u << slot;
        }
        __syncwarp();
        if (lane == 0)
            bulk_store(reinterpret_cast<char*>(out) + (size_t)tile * kTmaRows * row_bytes,
                       smem_addr(my_smem + (size_t)slot * stage_bytes), (uint32_t)rows * row_bytes);
    };

    while (drain_tile < num_tiles) {
        while (in_flight < kTmaStages - 1 && issue_tile < num_tiles) {
            issue(issue_tile, issue_slot);
            issue_tile += nwarps;
            issue_slot = issue_slot + 1 == kTmaStages ? 0 : issue_slot + 1;
            ++in_flight;
        }
        drain(drain_tile, drain_slot);
        drain_tile += nwarps;
        drain_slot = drain_slot + 1 == kTmaStages ? 0 : drain_slot + 1;
        --in_flight;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace

// true when the TMA gather handles this call (otherwise the caller launches the LDG/STG kernel)
bool bag_forward_tma_launch(const cebag_bag_args* a, const BagParams& p, float* out, cudaStream_t stream, int* rc) {
    // OFF by default (CEBAG_FWD_TMA=1 turns it on): measured 288 us vs 186 us for the LDG.128 / STG.128 kernel at
    // Criteo-1TB (1.7 M rows of 512 B per launch).  Every row is one bulk-copy operation, and an SM's TMA unit retires
    // roughly one operation per ~46 cycles (B300_MICROARCH.md "TMA service/SM"): 1.7 M ops x 46 cyc / 148 SMs / 1.965 GHz
    // = 270 us -- the gather is bound by the TMA issue rate, not by bytes; bulk copies need >= ~1 KB per operation to reach
    // HBM speed (profiles/r2_forward_tma.summary.txt).  Kept for wide rows and as the record of the experiment.
    static const int enabled = env_int("CEBAG_FWD_TMA", 0);
    static const int ctas_per_sm = env_int("CEBAG_FWD_TMA_CTAS_PER_SM", 2);
    *rc = CEBAG_OK;
    const int row_bytes = a->dim * (int)sizeof(float);
    if (!enabled || a->per_sample_weights || a->mode != CEBAG_MODE_SUM || a->layout != CEBAG_LAYOUT_BAG_MAJOR) return false;
    if (row_bytes % 16 != 0 || row_bytes < 64 || row_bytes > 1024) return false;
    if (!aligned16(a->cache) || !aligned16(out)) return false;
    if (a->num_bags < 4096) return false;                     // tiny calls: not worth the pipeline prologue
    const size_t smem = (size_t)kTmaWarps * kTmaStages * kTmaRows * row_bytes;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(bag_forward_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 / 2);
        if (e != cudaSuccess) { cudaGetLastError(); return false; }
        configured = true;
    }
    if (smem > 227 * 1024 / 2) return false;
    const int64_t tiles = ceil_div(a->num_bags, kTmaRows);
    int64_t grid = ceil_div(tiles, kTmaWarps);
    if (grid > (int64_t)kNumSMs * ctas_per_sm) grid = (int64_t)kNumSMs * ctas_per_sm;
    KernelScope scope(kKernForward, stream);
    bag_forward_tma_kernel<<<(int)grid, kTmaWarps * 32, smem, stream>>>(p, out, row_bytes);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("bag_forward_tma launch failed: %s", cudaGetErrorString(e));
        *rc = CEBAG_ERR_CUDA;
    }
    return true;
}

}  // namespace cebag
