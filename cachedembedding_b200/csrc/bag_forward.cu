// Offsets-based segment-sum gather over the slot cache: the forward of F.embedding_bag on cuda_cached_weight
// (reference call site recsys/models/dlrm.py:99-110).
//
// HBM/L2-bound gather: per lookup 8 B slot id + 4D B row, per bag 4D B output (+ offsets).  A group of LANES threads
// (row width / 16 B) works on LANES consecutive bags per iteration; every lane moves 128-bit chunks.
#include "bag_common.cuh"
#include "profile.cuh"

namespace cebag {

namespace {

constexpr int kFwdThreads = 256;

// lanes of this thread's group inside its warp (groups of one warp may be at different points of their loops)
template <int LANES>
__device__ __forceinline__ unsigned group_mask() {
    return LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (lane_id() & ~(LANES - 1)));
}

// A group of LANES threads takes LANES consecutive bags per iteration: lane l reads the offsets, the first slot id and
// the first weight of bag g0 + l with coalesced loads (with pooling factor 1 that is all there is to read), then the
// group walks the LANES bags, kUnroll at a time: slot ids are broadcast by shuffle, the kUnroll row loads are issued
// back to back, and only bags longer than one entry enter the per-entry loop.  ~12 instructions per row instead of
// >100 for the one-bag-per-group formulation (ncu: that one was issue-bound at 25 % occupancy).
// FAST: mode sum, no per-sample weights, no padding index, bag-major output -- the reference's DLRM call; the weight,
// count, padding and output-pointer bookkeeping compiles away.
// LDP: L1 policy of the row loads (Vec::ld_row).  Ids arrive feature-major and the grid walks the bags in order, so at
// any moment every SM works on the same table: the few rows of a small table are read by all warps of all SMs at the
// same time.  Allocating them in L1 (LDP = 1) takes those reads off the L2 slices that hold the rows: 188 -> 165 us at
// Criteo-1TB (L1 hit rate 0.8 % -> 35 %); bypassing L1 (LDP = 0, "every row is touched once") was round 1's choice.
template <typename VT, int LANES, int CPL, int kUnroll, bool FAST, int LDP>
__global__ void __launch_bounds__(kFwdThreads)
bag_forward_kernel(const BagParams p, float* __restrict__ out) {
    const int lane = threadIdx.x & (LANES - 1);
    const unsigned gmask = group_mask<LANES>();
    // groups are independent: the CTA size is a launch parameter (blockDim.x = 128 or 256)
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int64_t num_groups = (int64_t)gridDim.x * blockDim.x / LANES;
    const VT* __restrict__ cache = reinterpret_cast<const VT*>(p.cache);
    const int chunks = p.chunks;
    const bool mean = !FAST && p.mode == CEBAG_MODE_MEAN;
    const float* __restrict__ psw = FAST ? nullptr : p.psw;
    const long long padding = FAST ? -1 : (long long)p.padding_idx;

    for (int64_t g0 = group * LANES; g0 < p.num_bags; g0 += num_groups * LANES) {
        const int64_t g = g0 + lane;
        const bool have = g < p.num_bags;
        const int64_t my_lo = have ? load_offset(p, g) : 0;
        const int64_t my_hi = have ? load_offset(p, g + 1) : 0;
        const int my_len = (int)(my_hi - my_lo);
        long long my_slot = -1;
        float my_w = 1.f;
        if (my_len > 0) {
            my_slot = __ldg(p.slot_ids + my_lo);
            if (psw) my_w = __ldg(psw + my_lo);
            if (my_slot == padding) my_slot = -1;
        }
        const float* my_out = FAST ? nullptr : bag_row_ptr(p, out, have ? g : 0);
        const int nb = (int)min((int64_t)LANES, p.num_bags - g0);

#pragma unroll 1
        for (int u0 = 0; u0 < nb; u0 += kUnroll) {
            int sl[kUnroll], len[kUnroll];
            float w[kUnroll];
            VT acc[kUnroll][CPL];
#pragma unroll
            for (int k = 0; k < kUnroll; ++k) {
                const int src = min(u0 + k, LANES - 1);
                sl[k] = (int)__shfl_sync(gmask, (int)my_slot, src, LANES);       // slot ids < 2^31
                len[k] = __shfl_sync(gmask, my_len, src, LANES);
                w[k] = FAST ? 1.f : __shfl_sync(gmask, my_w, src, LANES);
                if (u0 + k >= nb) { sl[k] = -1; len[k] = 0; }
            }
#pragma unroll
            for (int k = 0; k < kUnroll; ++k) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int col = lane + c * LANES;
                    acc[k][c] = (sl[k] >= 0 && col < chunks)
                                    ? Vec<VT>::template ld_row<LDP>(cache + (int64_t)sl[k] * chunks + col) : Vec<VT>::zero();
                }
            }
#pragma unroll
            for (int k = 0; k < kUnroll; ++k) {
                if (u0 + k >= nb) continue;
                int cnt = sl[k] >= 0 ? 1 : 0;
                if (psw) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) acc[k][c] = Vec<VT>::scale(acc[k][c], w[k]);
                }
                if (len[k] > 1) {        // group-uniform: bags with more than one entry
                    const int src = u0 + k;
                    const int64_t lo = ((int64_t)__shfl_sync(gmask, (int)(my_lo >> 32), src, LANES) << 32) |
                                       (unsigned)__shfl_sync(gmask, (int)(my_lo & 0xffffffff), src, LANES);
                    for (int i = 1; i < len[k]; ++i) {
                        long long s2 = __ldg(p.slot_ids + lo + i);
                        if (s2 == padding) continue;
                        const float w2 = psw ? __ldg(psw + lo + i) : 1.f;
                        ++cnt;
#pragma unroll
                        for (int c = 0; c < CPL; ++c) {
                            const int col = lane + c * LANES;
                            if (col < chunks)
                                Vec<VT>::fma(acc[k][c], w2, Vec<VT>::template ld_row<LDP>(cache + s2 * chunks + col));
                        }
                    }
                }
                VT* orow = FAST ? reinterpret_cast<VT*>(out) + (g0 + u0 + k) * chunks
                                : reinterpret_cast<VT*>(const_cast<float*>(shfl_ptr(gmask, my_out, u0 + k, LANES)));
                const float scale = (mean && cnt > 0) ? 1.f / (float)cnt : 1.f;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int col = lane + c * LANES;
                    if (col < chunks) {
                        VT r = mean ? Vec<VT>::scale(acc[k][c], scale) : acc[k][c];
                        Vec<VT>::st_stream(orow + col, r);
                    }
                }
            }
        }
    }
}

}  // namespace

int fill_bag_params(const cebag_bag_args* a, BagParams* p, const RowShape& rs) {
    CEBAG_REQUIRE(a != nullptr, "null args");
    CEBAG_REQUIRE(a->dim > 0 && a->cache_rows > 0, "cache shape");
    CEBAG_REQUIRE(a->n >= 0 && a->num_bags >= 0 && a->num_bags < ((int64_t)1 << 31), "sizes");
    CEBAG_REQUIRE(a->cache_rows > 0 && (int64_t)a->cache_rows < ((int64_t)1 << 31), "cache_rows");
    CEBAG_REQUIRE(a->n == 0 || a->slot_ids != nullptr, "slot_ids");
    CEBAG_REQUIRE(a->num_bags == 0 || a->offsets != nullptr, "offsets");
    CEBAG_REQUIRE(a->mode == CEBAG_MODE_SUM || a->mode == CEBAG_MODE_MEAN, "mode");
    CEBAG_REQUIRE(!(a->per_sample_weights && a->mode != CEBAG_MODE_SUM), "per_sample_weights need mode sum");
    p->cache = a->cache;
    p->slot_ids = a->slot_ids;
    p->offsets = a->offsets;
    p->psw = a->per_sample_weights;
    p->n = a->n;
    p->num_bags = a->num_bags;
    p->padding_idx = a->padding_idx >= 0 ? a->padding_idx : -1;
    p->dim = a->dim;
    p->cache_rows = a->cache_rows;
    p->key_mask = 0xffffffffu;
    p->chunks = rs.chunks;
    p->offsets_are_64 = a->offsets_are_64;
    p->include_last = a->include_last_offset;
    p->mode = a->mode;
    p->layout = a->layout;
    p->layout_batch = 1;
    p->layout_features = 1;
    p->exch_world = 0;
    if (a->layout == CEBAG_LAYOUT_SAMPLE_MAJOR) {
        CEBAG_REQUIRE(a->layout_batch > 0 && a->num_bags % a->layout_batch == 0, "sample-major layout needs G = F * B");
        p->layout_batch = a->layout_batch;
        p->layout_features = a->num_bags / a->layout_batch;
    } else if (a->layout == CEBAG_LAYOUT_EXCHANGE) {
        const cebag_exchange* x = a->exchange;
        CEBAG_REQUIRE(x != nullptr, "exchange layout needs a cebag_exchange");
        CEBAG_REQUIRE(x->world >= 1 && x->world <= CEBAG_MAX_PEERS, "exchange world");
        CEBAG_REQUIRE(a->layout_batch >= x->world && a->num_bags % a->layout_batch == 0, "exchange layout needs G = F_local * B");
        p->layout_batch = a->layout_batch;
        p->layout_features = a->num_bags / a->layout_batch;
        CEBAG_REQUIRE(x->feature_offset >= 0 && x->feature_offset + p->layout_features <= x->total_features,
                      "exchange feature range");
        p->exch_world = x->world;
        p->exch_feature_offset = x->feature_offset;
        p->exch_total_features = x->total_features;
        p->exch_base = (int32_t)(a->layout_batch / x->world);
        p->exch_rem = (int32_t)(a->layout_batch % x->world);
        for (int q = 0; q < CEBAG_MAX_PEERS; ++q) {
            p->exch_peer[q] = q < x->world ? x->peer[q] : nullptr;
            CEBAG_REQUIRE(q >= x->world || (x->peer[q] != nullptr && aligned16(x->peer[q])), "exchange peer pointer");
        }
    } else {
        CEBAG_REQUIRE(a->layout == CEBAG_LAYOUT_BAG_MAJOR, "layout");
    }
    return CEBAG_OK;
}

}  // namespace cebag

using namespace cebag;

extern "C" int cebag_bag_forward(const cebag_bag_args* a, float* out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(a != nullptr, "null args");
    if (a->num_bags == 0) return CEBAG_OK;
    CEBAG_REQUIRE(out != nullptr || a->layout == CEBAG_LAYOUT_EXCHANGE, "out");
    RowShape rs = row_shape(a->dim, aligned16(a->cache) && aligned16(out));
    BagParams p;
    int rc = fill_bag_params(a, &p, rs);
    if (rc) return rc;
    {   // pooling-factor-1 style calls: TMA row gather (bulk async copies through shared memory)
        int tma_rc = CEBAG_OK;
        if (bag_forward_tma_launch(a, p, out, stream, &tma_rc)) return tma_rc;
    }
    // tuning knobs, read per call (an in-process sweep can compare them): rows in flight per group (4 | 8), CTAs per SM
    // of the grid-stride launch (32: 6.4 waves of the 5 resident CTAs per SM, 157 vs 164 us with 16), L1 policy of the
    // row loads (0 bypass, 1 allocate, 2 allocate + evict last)
    const int unroll_env = env_int("CEBAG_FWD_UNROLL", 4);
    const int ctas_per_sm = env_int("CEBAG_FWD_CTAS_PER_SM", 32);
    const int ld_policy = env_int("CEBAG_FWD_LD", 1);
    const int fwd_threads = env_int("CEBAG_FWD_THREADS", 256) == 128 ? 128 : 256;     // CTA size
    const bool fast_path = a->per_sample_weights == nullptr && a->mode == CEBAG_MODE_SUM && a->padding_idx < 0 &&
                           a->layout == CEBAG_LAYOUT_BAG_MAJOR;
#define LAUNCH_FWD_U(VT, LANES, CPL, UNROLL)                                                             \
    do {                                                                                                 \
        int64_t groups = ceil_div(p.num_bags, LANES);                                                    \
        int grid = grid_for(groups * LANES, fwd_threads, ctas_per_sm * (kFwdThreads / fwd_threads));     \
        if (fast_path && ld_policy == 0)                                                                 \
            bag_forward_kernel<VT, LANES, CPL, UNROLL, true, 0><<<grid, fwd_threads, 0, stream>>>(p, out);          \
        else if (fast_path && ld_policy == 2)                                                            \
            bag_forward_kernel<VT, LANES, CPL, UNROLL, true, 2><<<grid, fwd_threads, 0, stream>>>(p, out);          \
        else if (fast_path)                                                                              \
            bag_forward_kernel<VT, LANES, CPL, UNROLL, true, 1><<<grid, fwd_threads, 0, stream>>>(p, out);          \
        else bag_forward_kernel<VT, LANES, CPL, UNROLL, false, 1><<<grid, fwd_threads, 0, stream>>>(p, out);        \
    } while (0)
#define LAUNCH_FWD(VT, LANES, CPL)                                                                       \
    do {                                                                                                 \
        if (unroll_env >= 8 && CPL == 1 && LANES >= 8) LAUNCH_FWD_U(VT, LANES, CPL, 8);                  \
        else if (CPL <= 2) LAUNCH_FWD_U(VT, LANES, CPL, 4);                                              \
        else LAUNCH_FWD_U(VT, LANES, CPL, 2);                                                            \
    } while (0)
    KernelScope scope(kKernForward, stream);
    CEBAG_DISPATCH_ROW_SHAPE(rs, LAUNCH_FWD);
#undef LAUNCH_FWD
#undef LAUNCH_FWD_U
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}
