// Backward of the embedding bag over the slot cache, fused with the optimizer step on the cached rows
// (SURVEY.md K12 + K13; replaces _embedding_bag_sparse_backward + coalesce + torch.optim.SGD's sparse branch,
// reference call sites recsys/dlrm_main.py:274-279).
//
// Plan (deterministic, no float atomics):
//   1. bag_of[i]  : lookup position -> bag                                   (4 B/lookup)
//   2. radix sort : (slot, value) by slot, stable; value = bag (fast path) or lookup position
//   3. phase 1    : one group of LANES threads per chunk of kChunk sorted positions walks its runs of equal
//                   slots, summing w_i * grad_out[bag(i)] in index order.  Runs that lie inside the chunk are
//                   applied to the cached row at once (row read issued together with the grad reads); the
//                   first/last run of a chunk that continue into a neighbour are parked as partial sums.
//   4. phase 2    : one group per run that started in a chunk and crossed its end: adds the parked partials of
//                   the following chunks in order and applies the update.
// Algorithmic HBM bytes: 4D per lookup (grad row) + 8D per unique slot (row read + write) + sort traffic
// (~40 B/lookup).  The same machinery writes a dense [C, D] grad (sparse=False) instead of updating.
#include "bag_common.cuh"
#include "radix_sort.cuh"
#include "profile.cuh"

namespace cebag {

namespace {

constexpr int kBwdThreads = 256;
#ifndef CEBAG_BWD_CHUNK
#define CEBAG_BWD_CHUNK 64
#endif
constexpr int kChunk = CEBAG_BWD_CHUNK;   // sorted positions per group
constexpr int kSuperChunks = 16;  // chunks per super-chunk (phase 2a)

enum : int { kOptSgd = 0, kOptAdagrad = 1, kOptDense = 2 };
enum : unsigned char { kFlagOpenLeft = 1, kFlagOpenRight = 2, kFlagWhole = 4 };

__global__ void __launch_bounds__(256) bag_of_kernel(const BagParams p, int32_t* __restrict__ bag_of) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; g < p.num_bags; g += stride) {
        int64_t lo = load_offset(p, g), hi = load_offset(p, g + 1);
        for (int64_t i = lo; i < hi; ++i) bag_of[i] = (int32_t)g;
    }
}

// window plan: the pairs of batch `batch` -- key = (batch << slot_bits) | slot, value = bag -- written where the one
// radix sort of the whole window reads them
__global__ void __launch_bounds__(256)
window_pairs_kernel(const BagParams p, uint32_t batch, int slot_bits, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t mask = (1u << slot_bits) - 1u;
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; g < p.num_bags; g += stride) {
        const int64_t lo = load_offset(p, g), hi = load_offset(p, g + 1);
        for (int64_t i = lo; i < hi; ++i) {
            keys[i] = (batch << slot_bits) | ((uint32_t)p.slot_ids[i] & mask);
            vals[i] = (uint32_t)g;
        }
    }
}

// per-lookup weight for the slow path: psw[i] (sum) or 1 / (#non-padding entries of the bag) (mean)
__global__ void __launch_bounds__(256) lookup_weight_kernel(const BagParams p, float* __restrict__ wts) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; g < p.num_bags; g += stride) {
        int64_t lo = load_offset(p, g), hi = load_offset(p, g + 1);
        float scale = 1.f;
        if (p.mode == CEBAG_MODE_MEAN) {
            int32_t c = 0;
            for (int64_t i = lo; i < hi; ++i) c += (p.slot_ids[i] != p.padding_idx);
            scale = c > 0 ? 1.f / (float)c : 0.f;
        }
        for (int64_t i = lo; i < hi; ++i) wts[i] = p.psw ? p.psw[i] * scale : scale;
    }
}

struct UpdateParams {
    float* cache;          // fp32[C, D] updated in place (or dense grad target for kOptDense)
    float* state;          // fp32[C] row-wise Adagrad accumulators
    float  lr, eps;
    int32_t dim;
};

// apply the accumulated grad `acc` of slot `slot`; `w` holds the current row (prefetched) unless OPT == dense
template <typename VT, int LANES, int CPL, int OPT>
__device__ __forceinline__ void apply_update(const UpdateParams& up, int chunks, int lane, uint32_t slot,
                                             const VT (&acc)[CPL], const VT (&w)[CPL]) {
    VT* row = reinterpret_cast<VT*>(up.cache) + (int64_t)slot * chunks;
    if (OPT == kOptDense) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            int col = lane + c * LANES;
            if (col < chunks) Vec<VT>::st(row + col, acc[c]);
        }
        return;
    }
    float step = up.lr;
    if (OPT == kOptAdagrad) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            int col = lane + c * LANES;
            if (col < chunks) ss += Vec<VT>::dot(acc[c], acc[c]);
        }
        ss = group_sum<LANES>(ss);
        float st = up.state[slot] + ss / (float)up.dim;
        if (lane == 0) up.state[slot] = st;
        step = up.lr / (sqrtf(st) + up.eps);
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        int col = lane + c * LANES;
        if (col < chunks) Vec<VT>::st(row + col, Vec<VT>::sub_scaled(w[c], step, acc[c]));
    }
}

template <typename VT, int LANES, int CPL>
__device__ __forceinline__ void store_partial(float* scratch, int64_t chunk, int which, int chunks, int lane,
                                              const VT (&acc)[CPL]) {
    VT* dst = reinterpret_cast<VT*>(scratch) + (chunk * 2 + which) * chunks;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        int col = lane + c * LANES;
        if (col < chunks) Vec<VT>::st(dst + col, acc[c]);
    }
}

// Adds the parked first-run partials of elements h, h+1, ... (chunks or super-chunks) to `acc`, in order, while each
// element is flagged Whole (one run covers it entirely and continues); the first element that is not Whole is added
// too and closes the run.  kBatch loads are in flight at a time.  Returns true when the run closed inside [h, hend).
template <typename VT, int LANES, int CPL>
__device__ __forceinline__ bool chain_sum(const VT* __restrict__ sv, const unsigned char* __restrict__ flags,
                                          int64_t h, int64_t hend, int chunks, int lane, VT (&acc)[CPL]) {
    constexpr int kBatch = 8;
    bool ended = false;
    while (!ended && h < hend) {
        int take = 0;
        while (take < kBatch && h + take < hend) {
            unsigned char f = flags[h + take];
            ++take;
            if (!(f & kFlagWhole)) { ended = true; break; }
        }
        VT part[kBatch][CPL];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                int col = lane + c * LANES;
                part[b][c] = (b < take && col < chunks) ? Vec<VT>::ld(sv + ((h + b) * 2) * chunks + col) : Vec<VT>::zero();
            }
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            if (b < take) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) Vec<VT>::fma(acc[c], 1.f, part[b][c]);
            }
        }
        h += take;
    }
    return ended;
}

template <typename VT, int LANES, int CPL, int OPT>
__device__ __forceinline__ void load_row_and_apply(const UpdateParams& up, int chunks, int lane, uint32_t slot,
                                                   const VT (&acc)[CPL]) {
    VT wr[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        int col = lane + c * LANES;
        wr[c] = (col < chunks && OPT != kOptDense)
                    ? Vec<VT>::ld(reinterpret_cast<const VT*>(up.cache) + (int64_t)slot * chunks + col)
                    : Vec<VT>::zero();
    }
    apply_update<VT, LANES, CPL, OPT>(up, chunks, lane, slot, acc, wr);
}

// Phase 1: one group per chunk of kChunk sorted positions; groups are independent (no block-level barrier).
// The group reads its (slot, value) pairs LANES at a time with coalesced loads (lane l owns position sb + l and also
// looks at the next key, which tells it whether a run ends there), then walks them kUnroll at a time with shuffles:
// the grad rows -- and the cached row wherever a run ends -- of kUnroll positions are loaded back to back.
// VAL_IS_BAG: sorted values are bag ids and every weight is 1 (mode sum, no per-sample weights)
template <typename VT, int LANES, int CPL, int OPT, bool VAL_IS_BAG, int kUnroll>
__global__ void __launch_bounds__(kBwdThreads)
bag_backward_phase1_kernel(const BagParams p, const UpdateParams up, const uint32_t* __restrict__ keys,
                           const uint32_t* __restrict__ vals, const int32_t* __restrict__ bag_of,
                           const float* __restrict__ wts, const float* __restrict__ grad_out,
                           float* __restrict__ scratch, unsigned char* __restrict__ flags, int64_t num_chunks) {
    const int lane = threadIdx.x & (LANES - 1);
    const unsigned gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (lane_id() & ~(LANES - 1)));
    const int gshift = lane_id() & ~(LANES - 1);
    // groups are independent: the CTA size is a launch parameter (blockDim.x = 128 or 256)
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int64_t num_groups = (int64_t)gridDim.x * blockDim.x / LANES;
    const int chunks = p.chunks;
    const VT* __restrict__ cachev = reinterpret_cast<const VT*>(up.cache);
    const uint32_t pad = p.padding_idx >= 0 ? (uint32_t)p.padding_idx : 0xffffffffu;
    const uint32_t kmask = p.key_mask, nslots = (uint32_t)p.cache_rows;
    // the slot of a sorted key; keys of padding entries and of invalid slot ids (-1) update nothing
    auto slot_of = [&](uint32_t key) { return key & kmask; };
    auto skip = [&](uint32_t slot) { return slot == pad || slot >= nslots; };

    for (int64_t ck = group; ck < num_chunks; ck += num_groups) {
        const int64_t start = ck * kChunk;
        const int64_t end = min(start + (int64_t)kChunk, p.n);
        bool first_run = true;
        const bool open_left = start > 0 && keys[start - 1] == keys[start];
        unsigned char flag = 0;
        VT acc[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[c] = Vec<VT>::zero();
        bool pending = false;

#pragma unroll 1
        for (int64_t sb = start; sb < end; sb += LANES) {
            const int64_t pos = sb + lane;
            const bool mine = pos < end;
            const uint32_t my_key = mine ? keys[pos] : 0u;
            const uint32_t my_next = (mine && pos + 1 < p.n) ? keys[pos + 1] : ~my_key;
            const uint32_t my_val = mine ? vals[pos] : 0u;
            const int my_bag = VAL_IS_BAG ? (int)my_val : (mine ? bag_of[my_val] : 0);
            const float my_w = (VAL_IS_BAG || !mine) ? 1.f : wts[my_val];
            const float* my_gptr = bag_row_ptr(p, const_cast<float*>(grad_out), my_bag);
            const unsigned tailbits = __ballot_sync(gmask, mine && my_next != my_key) >> gshift;
            const int nl = (int)min((int64_t)LANES, end - sb);

#pragma unroll 1
            for (int u0 = 0; u0 < nl; u0 += kUnroll) {
                uint32_t k[kUnroll];
                const VT* grow[kUnroll];
                float w[kUnroll];
                bool live[kUnroll], tail[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int src = min(u0 + u, LANES - 1);
                    live[u] = u0 + u < nl;
                    k[u] = __shfl_sync(gmask, my_key, src, LANES);
                    grow[u] = reinterpret_cast<const VT*>(shfl_ptr(gmask, my_gptr, src, LANES));
                    w[u] = VAL_IS_BAG ? 1.f : __shfl_sync(gmask, my_w, src, LANES);
                    tail[u] = live[u] && ((tailbits >> (u0 + u)) & 1u);
                }
                VT gr[kUnroll][CPL], wr[kUnroll][CPL];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        const int col = lane + c * LANES;
                        const bool ok = live[u] && col < chunks;
                        gr[u][c] = ok ? Vec<VT>::ld_stream(grow[u] + col) : Vec<VT>::zero();
                        // current row, needed where a run ends inside this chunk
                        const bool need_row = ok && tail[u] && OPT != kOptDense && !skip(slot_of(k[u]));
                        wr[u][c] = need_row ? Vec<VT>::ld(cachev + (int64_t)slot_of(k[u]) * chunks + col) : Vec<VT>::zero();
                    }
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    if (!live[u]) continue;
#pragma unroll
                    for (int c = 0; c < CPL; ++c) Vec<VT>::fma(acc[c], w[u], gr[u][c]);
                    pending = true;
                    if (tail[u]) {
                        if (first_run && open_left) {
                            store_partial<VT, LANES, CPL>(scratch, ck, 0, chunks, lane, acc);
                            flag |= kFlagOpenLeft;
                        } else if (!skip(slot_of(k[u]))) {
                            apply_update<VT, LANES, CPL, OPT>(up, chunks, lane, slot_of(k[u]), acc, wr[u]);
                        }
#pragma unroll
                        for (int c = 0; c < CPL; ++c) acc[c] = Vec<VT>::zero();
                        first_run = false;
                        pending = false;
                    }
                }
            }
        }
        if (pending) {  // the last run continues into the next chunk
            if (first_run && open_left) {
                store_partial<VT, LANES, CPL>(scratch, ck, 0, chunks, lane, acc);
                flag |= kFlagOpenLeft | kFlagWhole;
            } else {
                store_partial<VT, LANES, CPL>(scratch, ck, 1, chunks, lane, acc);
                flag |= kFlagOpenRight;
            }
        }
        if (lane == 0) flags[ck] = flag;
    }
}

// Phase 2a: one group per CHUNK (round 1 walked the 16 chunks of a super-chunk with one group: in regions of
// medium-hot slots every chunk ends in a run that crosses into the next one, and 16 dependent chains of loads made the
// kernel 32 us of pure latency).  The group of chunk g finishes the run that starts in g and crosses its right end --
// if it ends inside g's super-chunk -- or parks its partial for phase 2b; the group of a super-chunk's first chunk also
// collapses the run that enters the super-chunk from the left into ONE partial, so that a slot hit by all 65536
// lookups of a tiny table is a chain of 64 partials in phase 2b, not 1024.  flags_left[S]: OpenLeft / Whole of the
// super-chunk (written by its first chunk's group); flags_right[S]: a run that started inside S crosses its right end
// (written by its last chunk's group).
template <typename VT, int LANES, int CPL, int OPT>
__global__ void __launch_bounds__(kBwdThreads)
bag_backward_phase2a_kernel(const BagParams p, const UpdateParams up, const uint32_t* __restrict__ keys,
                            const float* __restrict__ scratch_c, const unsigned char* __restrict__ flags_c,
                            int64_t num_chunks, float* __restrict__ scratch_s, unsigned char* __restrict__ flags_left,
                            unsigned char* __restrict__ flags_right, int64_t num_super) {
    const int lane = threadIdx.x & (LANES - 1);
    const int64_t group = ((int64_t)blockIdx.x * kBwdThreads + threadIdx.x) / LANES;
    const int64_t num_groups = (int64_t)gridDim.x * kBwdThreads / LANES;
    const int chunks = p.chunks;
    const VT* __restrict__ sv = reinterpret_cast<const VT*>(scratch_c);
    const uint32_t pad = p.padding_idx >= 0 ? (uint32_t)p.padding_idx : 0xffffffffu;
    static_assert(kSuperChunks == 16, "one uint4 of flags per super-chunk");
    for (int64_t g = group; g < num_chunks; g += num_groups) {
        const int64_t S = g / kSuperChunks;
        const int64_t c0 = S * kSuperChunks;
        const int64_t cend = min(c0 + (int64_t)kSuperChunks, num_chunks);
        const unsigned char f = flags_c[g];
        if (g == c0) {                                       // the run that enters this super-chunk from the left
            unsigned char sflag = 0;
            if (f & kFlagOpenLeft) {
                VT acc[CPL];
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    int col = lane + c * LANES;
                    acc[c] = col < chunks ? Vec<VT>::ld(sv + (c0 * 2) * chunks + col) : Vec<VT>::zero();
                }
                sflag |= kFlagOpenLeft;
                if (f & kFlagWhole) {
                    bool ended = chain_sum<VT, LANES, CPL>(sv, flags_c, c0 + 1, cend, chunks, lane, acc);
                    if (!ended) sflag |= kFlagWhole;         // one run covers the whole super-chunk and goes on
                }
                store_partial<VT, LANES, CPL>(scratch_s, S, 0, chunks, lane, acc);
            }
            if (lane == 0) flags_left[S] = sflag;
        }
        if (g == cend - 1 && lane == 0) {
            // does a run that STARTED inside this super-chunk cross its right end?  The last chunk tells whether any run
            // crosses (its last run is open to the right); it started inside unless every chunk is one and the same run
            bool crossed = cend < num_chunks && (f & (kFlagOpenRight | kFlagWhole));
            if (crossed && !(f & kFlagOpenRight)) {
                bool all_whole = (flags_c[c0] & kFlagOpenLeft) != 0;
                for (int64_t h = c0; h < cend && all_whole; ++h) all_whole = (flags_c[h] & kFlagWhole) != 0;
                crossed = !all_whole;
            }
            flags_right[S] = crossed ? 1 : 0;
        }
        if (f & kFlagOpenRight) {                            // chunk g holds the head of a run that crosses its end
            VT acc[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                int col = lane + c * LANES;
                acc[c] = col < chunks ? Vec<VT>::ld(sv + (g * 2 + 1) * chunks + col) : Vec<VT>::zero();
            }
            bool ended = chain_sum<VT, LANES, CPL>(sv, flags_c, g + 1, cend, chunks, lane, acc);
            if (ended) {
                const uint32_t slot = keys[min((g + 1) * (int64_t)kChunk, p.n) - 1] & p.key_mask;
                if (slot != pad && slot < (uint32_t)p.cache_rows)
                    load_row_and_apply<VT, LANES, CPL, OPT>(up, chunks, lane, slot, acc);
            } else {                                         // continues into the next super-chunk
                store_partial<VT, LANES, CPL>(scratch_s, S, 1, chunks, lane, acc);
            }
        }
    }
}

// Phase 2b: one CTA per run that started inside a super-chunk and crossed its end (a hot slot of a tiny table: up to 64
// super-chunk partials).  Round 1 walked such a chain with one group, 8 loads at a time -- ~25 us of serial latency for
// a few dozen runs.  Here the CTA's groups each take 8 consecutive partials (all loaded at once), sum them in order up
// to the end of the run, and group 0 adds the groups' sums in order: the same result on every run, one load latency.
template <typename VT, int LANES, int CPL, int OPT>
__global__ void __launch_bounds__(kBwdThreads)
bag_backward_phase2b_kernel(const BagParams p, const UpdateParams up, const uint32_t* __restrict__ keys,
                            const float* __restrict__ scratch_s, const unsigned char* __restrict__ flags_left,
                            const unsigned char* __restrict__ flags_right, int64_t num_super) {
    constexpr int64_t kSuper = (int64_t)kSuperChunks * kChunk;
    constexpr int kGroups = kBwdThreads / LANES;
    constexpr int kPer = 8;
    __shared__ VT part[kGroups][LANES * CPL];
    __shared__ int take_s[kGroups], ended_s[kGroups], done_s;
    const int lane = threadIdx.x & (LANES - 1);
    const int gi = threadIdx.x / LANES;
    const int chunks = p.chunks;
    const VT* __restrict__ sv = reinterpret_cast<const VT*>(scratch_s);
    const uint32_t pad = p.padding_idx >= 0 ? (uint32_t)p.padding_idx : 0xffffffffu;
    for (int64_t S = blockIdx.x; S < num_super; S += gridDim.x) {
        if (!flags_right[S]) continue;                      // uniform over the CTA
        const uint32_t slot = keys[min((S + 1) * kSuper, p.n) - 1] & p.key_mask;
        VT total[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            int col = lane + c * LANES;
            total[c] = (gi == 0 && col < chunks) ? Vec<VT>::ld(sv + (S * 2 + 1) * chunks + col) : Vec<VT>::zero();
        }
        int64_t h = S + 1;
        bool done = h >= num_super;
        while (!done) {
            const int64_t base = h + (int64_t)gi * kPer;
            int take = 0;
            bool ended = false;
            for (int b = 0; b < kPer && base + b < num_super; ++b) {
                ++take;
                if (!(flags_left[base + b] & kFlagWhole)) { ended = true; break; }
            }
            VT rows[kPer][CPL];
#pragma unroll
            for (int b = 0; b < kPer; ++b) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    int col = lane + c * LANES;
                    rows[b][c] = (b < take && col < chunks) ? Vec<VT>::ld(sv + ((base + b) * 2) * chunks + col) : Vec<VT>::zero();
                }
            }
            VT gsum[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) gsum[c] = Vec<VT>::zero();
#pragma unroll
            for (int b = 0; b < kPer; ++b) {
                if (b < take) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) Vec<VT>::fma(gsum[c], 1.f, rows[b][c]);
                }
            }
#pragma unroll
            for (int c = 0; c < CPL; ++c) part[gi][lane + c * LANES] = gsum[c];
            if (lane == 0) { take_s[gi] = take; ended_s[gi] = ended ? 1 : 0; }
            __syncthreads();
            if (gi == 0) {
                bool fin = false;
                for (int q = 0; q < kGroups && !fin; ++q) {
                    if (take_s[q] == 0) { fin = true; break; }       // ran off the end of the array
#pragma unroll
                    for (int c = 0; c < CPL; ++c) Vec<VT>::fma(total[c], 1.f, part[q][lane + c * LANES]);
                    if (ended_s[q]) fin = true;
                }
                if (lane == 0) done_s = fin ? 1 : 0;
            }
            __syncthreads();
            done = done_s != 0;
            h += (int64_t)kGroups * kPer;
            if (h >= num_super) done = true;
            __syncthreads();
        }
        if (gi == 0 && slot != pad && slot < (uint32_t)p.cache_rows)
            load_row_and_apply<VT, LANES, CPL, OPT>(up, chunks, lane, slot, total);
    }
}

// ---- compatibility forms ------------------------------------------------------------------------------------------
// COO values: values[i] = w_i * grad_out[bag(i)]
template <typename VT, int LANES, int CPL>
__global__ void __launch_bounds__(kBwdThreads)
bag_backward_coo_kernel(const BagParams p, const float* __restrict__ grad_out, float* __restrict__ values) {
    const int lane = threadIdx.x & (LANES - 1);
    const int64_t group = ((int64_t)blockIdx.x * kBwdThreads + threadIdx.x) / LANES;
    const int64_t num_groups = (int64_t)gridDim.x * kBwdThreads / LANES;
    const int chunks = p.chunks;
    const VT* __restrict__ gradv = reinterpret_cast<const VT*>(grad_out);
    VT* __restrict__ valv = reinterpret_cast<VT*>(values);
    for (int64_t g = group; g < p.num_bags; g += num_groups) {
        int64_t lo = load_offset(p, g), hi = load_offset(p, g + 1);
        if (hi <= lo) continue;
        float scale = 1.f;
        if (p.mode == CEBAG_MODE_MEAN) {
            int32_t cnt = 0;
            for (int64_t i = lo; i < hi; ++i) cnt += (p.slot_ids[i] != p.padding_idx);
            scale = cnt > 0 ? 1.f / (float)cnt : 0.f;
        }
        int64_t row = bag_row(p, g);
        VT gv[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            int col = lane + c * LANES;
            gv[c] = col < chunks ? Vec<VT>::ld_stream(gradv + row * chunks + col) : Vec<VT>::zero();
        }
        for (int64_t i = lo; i < hi; ++i) {
            float w = (p.psw ? p.psw[i] : 1.f) * scale;
            if (p.slot_ids[i] == p.padding_idx) w = 0.f;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                int col = lane + c * LANES;
                if (col < chunks) Vec<VT>::st_stream(valv + i * chunks + col, Vec<VT>::scale(gv[c], w));
            }
        }
    }
}

// grad of per_sample_weights: gw[i] = <grad_out[bag(i)], cache[slot_i]>
template <typename VT, int LANES, int CPL>
__global__ void __launch_bounds__(kBwdThreads)
bag_backward_weights_kernel(const BagParams p, const float* __restrict__ grad_out, float* __restrict__ gw) {
    const int lane = threadIdx.x & (LANES - 1);
    const int64_t group = ((int64_t)blockIdx.x * kBwdThreads + threadIdx.x) / LANES;
    const int64_t num_groups = (int64_t)gridDim.x * kBwdThreads / LANES;
    const int chunks = p.chunks;
    const VT* __restrict__ gradv = reinterpret_cast<const VT*>(grad_out);
    const VT* __restrict__ cache = reinterpret_cast<const VT*>(p.cache);
    for (int64_t g = group; g < p.num_bags; g += num_groups) {
        int64_t lo = load_offset(p, g), hi = load_offset(p, g + 1);
        if (hi <= lo) continue;
        int64_t row = bag_row(p, g);
        VT gv[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            int col = lane + c * LANES;
            gv[c] = col < chunks ? Vec<VT>::ld_stream(gradv + row * chunks + col) : Vec<VT>::zero();
        }
        for (int64_t i = lo; i < hi; ++i) {
            int64_t s = p.slot_ids[i];
            float d = 0.f;
            if (s != p.padding_idx) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    int col = lane + c * LANES;
                    if (col < chunks) d += Vec<VT>::dot(gv[c], Vec<VT>::ld_stream(cache + s * chunks + col));
                }
            }
            d = group_sum<LANES>(d);
            if (lane == 0) gw[i] = d;
        }
    }
}

struct BwdLayout {
    size_t sort, bag_of, wts, scratch, flags, scratch_s, flags_s, flags_r, total;
    int64_t num_chunks, num_super;
};

BwdLayout bwd_layout(int64_t n, int dim) {
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    BwdLayout L;
    int64_t nn = n > 0 ? n : 1;
    L.num_chunks = ceil_div(nn, kChunk);
    L.num_super = ceil_div(L.num_chunks, kSuperChunks);
    size_t off = 0;
    L.sort = off; off += align(radix_sort_workspace_bytes(nn));
    L.bag_of = off; off += align((size_t)nn * 4);
    L.wts = off; off += align((size_t)nn * 4);
    L.scratch = off; off += align((size_t)L.num_chunks * 2 * dim * 4);
    L.flags = off; off += align((size_t)L.num_super * kSuperChunks);
    L.scratch_s = off; off += align((size_t)L.num_super * 2 * dim * 4);
    L.flags_s = off; off += align((size_t)L.num_super);
    L.flags_r = off; off += align((size_t)L.num_super);
    L.total = off;
    return L;
}

// bits that tell all slot ids [0, C) apart AND from the all-ones pattern of an invalid slot (-1), which the masked
// digits of the sort would otherwise alias with slot 2^bits - 1
int key_bits_for(int32_t cache_rows) {
    int bits = 1;
    while (((int64_t)1 << bits) < (int64_t)cache_rows + 1) ++bits;
    return bits;
}

// bag_of (+ per-lookup weights) and the radix sort: everything of the backward that does not need the gradient
int plan_sorted_backward(const cebag_bag_args* a, const BagParams& p, const BwdLayout& L, char* ws, cudaStream_t stream) {
    int32_t* bag_of = reinterpret_cast<int32_t*>(ws + L.bag_of);
    float* wts = reinterpret_cast<float*>(ws + L.wts);
    const bool fast = (a->mode == CEBAG_MODE_SUM && a->per_sample_weights == nullptr);
    {
        KernelScope scope(kKernBagOf, stream, fast ? 1 : 2);
        bag_of_kernel<<<grid_for(a->num_bags, 256, env_int("CEBAG_PLAN_CTAS_PER_SM", 8)), 256, 0, stream>>>(p, bag_of);
        CEBAG_LAUNCH_CHECK();
        if (!fast) {
            lookup_weight_kernel<<<grid_for(a->num_bags, 256, 8), 256, 0, stream>>>(p, wts);
            CEBAG_LAUNCH_CHECK();
        }
    }
    const uint32_t *keys = nullptr, *vals = nullptr;
    return radix_sort_slots(a->slot_ids, a->n, key_bits_for(a->cache_rows), ws + L.sort,
                            radix_sort_workspace_bytes(a->n), fast ? reinterpret_cast<const uint32_t*>(bag_of) : nullptr,
                            &keys, &vals, stream);
}

template <int OPT>
int run_sorted_backward(const cebag_bag_args* a, const float* grad_out, float* target, float* state, float lr,
                        float eps, void* workspace, size_t workspace_bytes, int has_plan, cudaStream_t stream) {
    if (a->n == 0) return CEBAG_OK;
    CEBAG_REQUIRE((grad_out != nullptr || a->layout == CEBAG_LAYOUT_EXCHANGE) && target != nullptr, "grad_out / target");
    CEBAG_REQUIRE(workspace != nullptr, "workspace");
    CEBAG_REQUIRE(a->n < ((int64_t)1 << 31) && a->num_bags < ((int64_t)1 << 31), "backward size");
    BwdLayout L = bwd_layout(a->n, a->dim);
    CEBAG_REQUIRE(workspace_bytes >= L.total, "backward workspace too small");
    CEBAG_REQUIRE(aligned16(workspace), "workspace alignment");
    RowShape rs = row_shape(a->dim, aligned16(target) && aligned16(grad_out));
    BagParams p;
    int rc = fill_bag_params(a, &p, rs);
    if (rc) return rc;
    char* ws = reinterpret_cast<char*>(workspace);
    int32_t* bag_of = reinterpret_cast<int32_t*>(ws + L.bag_of);
    float* wts = reinterpret_cast<float*>(ws + L.wts);
    float* scratch = reinterpret_cast<float*>(ws + L.scratch);
    unsigned char* flags = reinterpret_cast<unsigned char*>(ws + L.flags);
    float* scratch_s = reinterpret_cast<float*>(ws + L.scratch_s);
    unsigned char* flags_s = reinterpret_cast<unsigned char*>(ws + L.flags_s);
    unsigned char* flags_r = reinterpret_cast<unsigned char*>(ws + L.flags_r);
    const bool fast = (a->mode == CEBAG_MODE_SUM && a->per_sample_weights == nullptr);
    CEBAG_REQUIRE(!has_plan || fast, "a backward plan exists only for mode sum without per-sample weights");
    CEBAG_REQUIRE(has_plan >= 0 && has_plan <= 2, "workspace_has_plan");
    const uint32_t *keys = nullptr, *vals = nullptr;
    if (has_plan == 2) {             // this batch's segment of a window plan
        CEBAG_REQUIRE(a->plan_keys && a->plan_vals && a->plan_key_mask, "window plan pointers");
        keys = a->plan_keys;
        vals = a->plan_vals;
        p.key_mask = a->plan_key_mask;
    } else {
        if (!has_plan) {
            rc = plan_sorted_backward(a, p, L, ws, stream);
            if (rc) return rc;
        }
        radix_sort_result(a->n, key_bits_for(a->cache_rows), ws + L.sort, &keys, &vals);
    }
    UpdateParams up;
    up.cache = target;
    up.state = state;
    up.lr = lr;
    up.eps = eps;
    up.dim = a->dim;
    // sorted positions in flight per group (tunable: CEBAG_BWD_UNROLL = 4 | 8; 8 only for one chunk per lane)
    const int unroll_env = env_int("CEBAG_BWD_UNROLL", 4);
    const int bwd_ctas_per_sm = env_int("CEBAG_BWD_CTAS_PER_SM", 32);
    // CTA size of phase 1 (128 | 256).  Smaller CTAs = finer granularity when a side-stream kernel takes registers on
    // the SM: with 74 registers only three 256-thread CTAs are resident, and a guest CTA evicts a third of them.
    const int bwd_threads = env_int("CEBAG_BWD_THREADS", 256) == 128 ? 128 : 256;
#define LAUNCH_P1(VT, LANES, CPL, FAST, UNROLL)                                                                     \
    bag_backward_phase1_kernel<VT, LANES, CPL, OPT, FAST, UNROLL><<<grid, bwd_threads, 0, stream>>>(                \
        p, up, keys, vals, bag_of, wts, grad_out, scratch, flags, L.num_chunks)
#define LAUNCH_BWD(VT, LANES, CPL)                                                                                  \
    do {                                                                                                            \
        {                                                                                                           \
            KernelScope scope1(kKernBwdPhase1, stream);                                                             \
            int grid = grid_for(L.num_chunks * LANES, bwd_threads, bwd_ctas_per_sm * (kBwdThreads / bwd_threads));  \
            if (fast) {                                                                                             \
                if (unroll_env >= 8 && CPL == 1) LAUNCH_P1(VT, LANES, CPL, true, 8);                                \
                else LAUNCH_P1(VT, LANES, CPL, true, 4);                                                            \
            } else {                                                                                                \
                LAUNCH_P1(VT, LANES, CPL, false, 4);                                                                \
            }                                                                                                       \
        }                                                                                                           \
        {                                                                                                           \
            KernelScope scope2(kKernBwdPhase2, stream, 2);                                                          \
            int grid_a = grid_for(L.num_chunks * LANES, kBwdThreads, 8);                                            \
            int grid_b = (int)(L.num_super < (int64_t)kNumSMs * 8 ? L.num_super : (int64_t)kNumSMs * 8);            \
            bag_backward_phase2a_kernel<VT, LANES, CPL, OPT><<<grid_a, kBwdThreads, 0, stream>>>(                   \
                p, up, keys, scratch, flags, L.num_chunks, scratch_s, flags_s, flags_r, L.num_super);               \
            bag_backward_phase2b_kernel<VT, LANES, CPL, OPT><<<grid_b, kBwdThreads, 0, stream>>>(                   \
                p, up, keys, scratch_s, flags_s, flags_r, L.num_super);                                             \
        }                                                                                                           \
    } while (0)
    CEBAG_DISPATCH_ROW_SHAPE(rs, LAUNCH_BWD);
#undef LAUNCH_BWD
#undef LAUNCH_P1
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}

}  // namespace
}  // namespace cebag

using namespace cebag;

extern "C" size_t cebag_backward_workspace_bytes(const cebag_bag_args* a) {
    if (!a) return 0;
    return bwd_layout(a->n, a->dim).total;
}

extern "C" int cebag_bag_backward_plan(const cebag_bag_args* a, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(a != nullptr, "null args");
    if (a->n == 0) return CEBAG_OK;
    CEBAG_REQUIRE(a->mode == CEBAG_MODE_SUM && a->per_sample_weights == nullptr,
                  "a backward plan exists only for mode sum without per-sample weights");
    CEBAG_REQUIRE(workspace != nullptr && aligned16(workspace), "workspace");
    CEBAG_REQUIRE(a->n < ((int64_t)1 << 31) && a->num_bags < ((int64_t)1 << 31), "backward size");
    BwdLayout L = bwd_layout(a->n, a->dim);
    CEBAG_REQUIRE(workspace_bytes >= L.total, "backward workspace too small");
    RowShape rs = row_shape(a->dim, true);
    BagParams p;
    int rc = fill_bag_params(a, &p, rs);
    if (rc) return rc;
    return plan_sorted_backward(a, p, L, reinterpret_cast<char*>(workspace), stream);
}

extern "C" size_t cebag_backward_window_plan_bytes(int64_t total_lookups) {
    return radix_sort_workspace_bytes(total_lookups > 0 ? total_lookups : 1);
}

extern "C" int cebag_bag_backward_plan_window(const cebag_bag_args* batches, int32_t num_batches, void* window_workspace,
                                              size_t workspace_bytes, const uint32_t** keys_out, const uint32_t** vals_out,
                                              uint32_t* mask_out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(batches != nullptr && num_batches >= 1 && num_batches <= 1024, "window batches");
    CEBAG_REQUIRE(keys_out && vals_out && mask_out, "window plan outputs");
    CEBAG_REQUIRE(window_workspace != nullptr && aligned16(window_workspace), "window workspace");
    int64_t total = 0;
    for (int j = 0; j < num_batches; ++j) {
        const cebag_bag_args* a = batches + j;
        CEBAG_REQUIRE(a->mode == CEBAG_MODE_SUM && a->per_sample_weights == nullptr,
                      "a backward plan exists only for mode sum without per-sample weights");
        CEBAG_REQUIRE(a->cache_rows == batches[0].cache_rows, "the batches of a window share one cache");
        CEBAG_REQUIRE(a->n >= 0 && a->num_bags < ((int64_t)1 << 31), "backward size");
        total += a->n;
    }
    CEBAG_REQUIRE(total < ((int64_t)1 << 30), "window size");
    const int slot_bits = key_bits_for(batches[0].cache_rows);
    int batch_bits = 0;
    while ((1 << batch_bits) < num_batches) ++batch_bits;
    CEBAG_REQUIRE(slot_bits + batch_bits <= 32, "window key width");
    *mask_out = (uint32_t)((1ull << slot_bits) - 1ull);
    if (total == 0) {
        for (int j = 0; j < num_batches; ++j) keys_out[j] = vals_out[j] = nullptr;
        return CEBAG_OK;
    }
    CEBAG_REQUIRE(workspace_bytes >= radix_sort_workspace_bytes(total), "window workspace too small");
    uint32_t *keys_in = nullptr, *vals_in = nullptr;
    radix_sort_input_buffers(total, window_workspace, &keys_in, &vals_in);
    int64_t begin = 0;
    {
        KernelScope scope(kKernBagOf, stream, num_batches);
        for (int j = 0; j < num_batches; ++j) {
            const cebag_bag_args* a = batches + j;
            if (a->n == 0 || a->num_bags == 0) continue;
            RowShape rs = row_shape(a->dim, true);
            BagParams p;
            int rc = fill_bag_params(a, &p, rs);
            if (rc) return rc;
            window_pairs_kernel<<<grid_for(a->num_bags, 256, 8), 256, 0, stream>>>(p, (uint32_t)j, slot_bits, keys_in + begin,
                                                                                 vals_in + begin);
            begin += a->n;
        }
        CEBAG_LAUNCH_CHECK();
    }
    const uint32_t *keys = nullptr, *vals = nullptr;
    int rc = radix_sort_u32(total, slot_bits + batch_bits, window_workspace, workspace_bytes, &keys, &vals, stream);
    if (rc) return rc;
    begin = 0;
    for (int j = 0; j < num_batches; ++j) {        // the batch index is the top of the key: segments are contiguous
        keys_out[j] = keys + begin;
        vals_out[j] = vals + begin;
        begin += batches[j].n;
    }
    return CEBAG_OK;
}

extern "C" int cebag_bag_backward_fused(const cebag_bag_args* a, const float* grad_out, float* cache_rw,
                                        float* cache_state, int32_t optimizer, float lr, float eps, void* workspace,
                                        size_t workspace_bytes, int32_t workspace_has_plan, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(a != nullptr, "null args");
    if (optimizer == CEBAG_OPT_SGD)
        return run_sorted_backward<kOptSgd>(a, grad_out, cache_rw, nullptr, lr, 0.f, workspace, workspace_bytes,
                                            workspace_has_plan, stream);
    if (optimizer == CEBAG_OPT_ROWWISE_ADAGRAD) {
        CEBAG_REQUIRE(cache_state != nullptr, "row-wise Adagrad needs cache_state");
        return run_sorted_backward<kOptAdagrad>(a, grad_out, cache_rw, cache_state, lr, eps, workspace,
                                                workspace_bytes, workspace_has_plan, stream);
    }
    set_error("unknown optimizer %d", optimizer);
    return CEBAG_ERR_INVALID;
}

extern "C" int cebag_bag_backward_dense(const cebag_bag_args* a, const float* grad_out, float* grad_cache,
                                        void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(a != nullptr && grad_cache != nullptr, "null args");
    CEBAG_CUDA_CHECK(cudaMemsetAsync(grad_cache, 0, (size_t)a->cache_rows * a->dim * sizeof(float), stream));
    return run_sorted_backward<kOptDense>(a, grad_out, grad_cache, nullptr, 0.f, 0.f, workspace, workspace_bytes,
                                          0, stream);
}

extern "C" int cebag_bag_backward_coo(const cebag_bag_args* a, const float* grad_out, float* values, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(a != nullptr, "null args");
    if (a->n == 0 || a->num_bags == 0) return CEBAG_OK;
    CEBAG_REQUIRE(grad_out != nullptr && values != nullptr, "grad_out / values");
    RowShape rs = row_shape(a->dim, aligned16(values) && aligned16(grad_out));
    BagParams p;
    int rc = fill_bag_params(a, &p, rs);
    if (rc) return rc;
#define LAUNCH_COO(VT, LANES, CPL)                                                                 \
    bag_backward_coo_kernel<VT, LANES, CPL><<<grid_for(p.num_bags * LANES, kBwdThreads, 8), kBwdThreads, 0, stream>>>( \
        p, grad_out, values)
    KernelScope scope(kKernBwdCoo, stream);
    CEBAG_DISPATCH_ROW_SHAPE(rs, LAUNCH_COO);
#undef LAUNCH_COO
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}

extern "C" int cebag_bag_backward_weights(const cebag_bag_args* a, const float* grad_out, float* grad_weights,
                                          void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(a != nullptr, "null args");
    if (a->n == 0 || a->num_bags == 0) return CEBAG_OK;
    CEBAG_REQUIRE(grad_out != nullptr && grad_weights != nullptr, "grad_out / grad_weights");
    RowShape rs = row_shape(a->dim, aligned16(a->cache) && aligned16(grad_out));
    BagParams p;
    int rc = fill_bag_params(a, &p, rs);
    if (rc) return rc;
#define LAUNCH_GW(VT, LANES, CPL)                                                                  \
    bag_backward_weights_kernel<VT, LANES, CPL><<<grid_for(p.num_bags * LANES, kBwdThreads, 8), kBwdThreads, 0, stream>>>( \
        p, grad_out, grad_weights)
    KernelScope scope(kKernBwdWeights, stream);
    CEBAG_DISPATCH_ROW_SHAPE(rs, LAUNCH_GW);
#undef LAUNCH_GW
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}
