// Offsets-based segment-sum gather over the slot cache: the forward of F.embedding_bag on cuda_cached_weight
// (reference call site recsys/models/dlrm.py:99-110).
//
// HBM-bound gather: per lookup 8 B slot id + 4D B row, per bag 4D B output (+ offsets).  One group of LANES
// threads owns a bag; every lane moves 128-bit chunks.  To keep enough bytes in flight per SM with pooling
// factor 1 (Criteo), each group works on kBagsPerIter consecutive bags at once: all offset loads, then all slot-id
// loads, then all row loads are issued before any is consumed.
#include "bag_common.cuh"
#include "profile.cuh"

namespace cebag {

namespace {

constexpr int kFwdThreads = 256;

template <typename VT, int LANES, int CPL, int kBagsPerIter>
__global__ void __launch_bounds__(kFwdThreads)
bag_forward_kernel(const BagParams p, float* __restrict__ out) {
    const int lane = threadIdx.x & (LANES - 1);
    const int64_t group = ((int64_t)blockIdx.x * kFwdThreads + threadIdx.x) / LANES;
    const int64_t num_groups = (int64_t)gridDim.x * kFwdThreads / LANES;
    const VT* __restrict__ cache = reinterpret_cast<const VT*>(p.cache);
    VT* __restrict__ outv = reinterpret_cast<VT*>(out);
    const int chunks = p.chunks;

    for (int64_t g0 = group * kBagsPerIter; g0 < p.num_bags; g0 += num_groups * kBagsPerIter) {
        int64_t lo[kBagsPerIter], hi[kBagsPerIter];
        int64_t prev = load_offset(p, g0);
        int64_t max_len = 0;
#pragma unroll
        for (int u = 0; u < kBagsPerIter; ++u) {
            bool valid = g0 + u < p.num_bags;
            int64_t next = valid ? load_offset(p, g0 + u + 1) : prev;
            lo[u] = prev;
            hi[u] = next;
            prev = next;
            max_len = max(max_len, hi[u] - lo[u]);
        }
        VT acc[kBagsPerIter][CPL];
        int32_t cnt[kBagsPerIter];
#pragma unroll
        for (int u = 0; u < kBagsPerIter; ++u) {
            cnt[u] = 0;
#pragma unroll
            for (int c = 0; c < CPL; ++c) acc[u][c] = Vec<VT>::zero();
        }
        // round t takes the t-th entry of each of the kBagsPerIter bags: independent loads across bags
        for (int64_t t = 0; t < max_len; ++t) {
            int64_t s[kBagsPerIter];
            float w[kBagsPerIter];
#pragma unroll
            for (int u = 0; u < kBagsPerIter; ++u) {
                int64_t i = lo[u] + t;
                bool live = i < hi[u];
                s[u] = live ? __ldg(p.slot_ids + i) : -1;
                w[u] = (live && p.psw) ? __ldg(p.psw + i) : 1.f;
                if (s[u] == p.padding_idx) s[u] = -1;
            }
            VT v[kBagsPerIter][CPL];
#pragma unroll
            for (int u = 0; u < kBagsPerIter; ++u) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    int col = lane + c * LANES;
                    v[u][c] = (s[u] >= 0 && col < chunks) ? Vec<VT>::ld_stream(cache + s[u] * chunks + col)
                                                          : Vec<VT>::zero();
                }
            }
#pragma unroll
            for (int u = 0; u < kBagsPerIter; ++u) {
                if (s[u] >= 0) {
                    cnt[u] += 1;
#pragma unroll
                    for (int c = 0; c < CPL; ++c) Vec<VT>::fma(acc[u][c], w[u], v[u][c]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kBagsPerIter; ++u) {
            if (g0 + u < p.num_bags) {
                float scale = (p.mode == CEBAG_MODE_MEAN && cnt[u] > 0) ? 1.f / (float)cnt[u] : 1.f;
                int64_t row = bag_row(p, g0 + u);
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    int col = lane + c * LANES;
                    if (col < chunks) {
                        VT r = p.mode == CEBAG_MODE_MEAN ? Vec<VT>::scale(acc[u][c], scale) : acc[u][c];
                        Vec<VT>::st_stream(outv + row * chunks + col, r);
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
    CEBAG_REQUIRE(a->n >= 0 && a->num_bags >= 0, "sizes");
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
    p->chunks = rs.chunks;
    p->offsets_are_64 = a->offsets_are_64;
    p->include_last = a->include_last_offset;
    p->mode = a->mode;
    p->layout = a->layout;
    p->layout_batch = 1;
    p->layout_features = 1;
    if (a->layout == CEBAG_LAYOUT_SAMPLE_MAJOR) {
        CEBAG_REQUIRE(a->layout_batch > 0 && a->num_bags % a->layout_batch == 0, "sample-major layout needs G = F * B");
        p->layout_batch = a->layout_batch;
        p->layout_features = a->num_bags / a->layout_batch;
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
    CEBAG_REQUIRE(out != nullptr, "out");
    RowShape rs = row_shape(a->dim, aligned16(a->cache) && aligned16(out));
    BagParams p;
    int rc = fill_bag_params(a, &p, rs);
    if (rc) return rc;
    // bags in flight per group: 8 / CPL by default (tunable: CEBAG_FWD_BPI = 2 | 4 | 8)
    static const int bpi_env = env_int("CEBAG_FWD_BPI", 0);
    static const int ctas_per_sm = env_int("CEBAG_FWD_CTAS_PER_SM", 8);
#define LAUNCH_FWD_BPI(VT, LANES, CPL, BPI)                                                              \
    do {                                                                                                 \
        int64_t groups = ceil_div(p.num_bags, BPI);                                                      \
        int grid = grid_for(groups * LANES, kFwdThreads, ctas_per_sm);                                   \
        bag_forward_kernel<VT, LANES, CPL, BPI><<<grid, kFwdThreads, 0, stream>>>(p, out);               \
    } while (0)
#define LAUNCH_FWD(VT, LANES, CPL)                                                                       \
    do {                                                                                                 \
        int bpi = bpi_env ? bpi_env : (CPL == 1 ? 8 : CPL == 2 ? 4 : 2);                                 \
        if (bpi >= 8 && CPL == 1) LAUNCH_FWD_BPI(VT, LANES, CPL, 8);                                     \
        else if (bpi >= 4 && CPL <= 2) LAUNCH_FWD_BPI(VT, LANES, CPL, 4);                                \
        else LAUNCH_FWD_BPI(VT, LANES, CPL, 2);                                                          \
    } while (0)
    KernelScope scope(kKernForward, stream);
    CEBAG_DISPATCH_ROW_SHAPE(rs, LAUNCH_FWD);
#undef LAUNCH_FWD
#undef LAUNCH_FWD_BPI
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}
