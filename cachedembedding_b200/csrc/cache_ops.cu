// The cache manager on the device: CachedParamMgr.prepare_ids / _prepare_rows_on_cuda / flush / reorder's preload
// (reference call site recsys/dlrm_main.py:259).
//
// The reference finds the unique rows of a window with sort-based unique/isin/topk over millions of ids, moves rows
// with CPU gather/scatter + staged memcpys and synchronises the device 8+ times per call.  Here one call ENQUEUES a
// fixed sequence of kernels and never waits for the GPU: every count (misses M, evictions E, free slots) stays in
// device memory, grids are upper bounds, and the state-changing kernels are gated by a device-side verdict.
//   * probe       : one pass over the ids.  Resident rows answer from the dense int32 row->slot map and raise their
//                   slot's hit flag (a plain byte store: no atomics on the hit path; the unique hits are counted from
//                   the flags afterwards); missing rows set a bit in a per-row bitmap and append their position to a
//                   fix-up list.  No sort.
//   * rank        : popcount-scan of the row bitmap emits the missed rows in ascending row order (the reference's
//                   sorted-unique contract) and their count M.
//   * decide      : one thread applies the A.3 capacity assert and the id-range check, fixes E = max(0, M - free).
//   * victims     : radix-select of the E smallest (freq, slot) [LFU] / largest row [DATASET] keys among slots that are
//                   neither hit by this call nor stamped by a protected window; ties by ascending slot.
//   * commit      : i-th smallest missed row -> i-th lowest free slot; maps, LFU counters, window stamps.
//   * row traffic : victims are ranked by HOST ROW through the same bitmap machinery -- a zero-copy scatter over PCIe
//                   runs at 48.8 GB/s when its destinations ascend and at 33.7 GB/s when they do not
//                   (scripts/probes/host_scatter_probe.cu) -- parked in an HBM staging buffer when a copy stream is
//                   used, the missed rows are gathered from the pinned host table (128-bit zero-copy loads, ascending
//                   rows), and the parked victims are then written back.  Reads and writes are separate kernels and
//                   never overlap: SM-issued PCIe reads and writes share one request queue (16 + 16 GB/s together).
//   * fix-up      : ids that missed get their new slot; LFU counters += multiplicity.
// Integer / byte work throughout; HBM- (maps) and PCIe- (rows) bound.
#include <string.h>
#include "common.cuh"
#include "scan.cuh"
#include "profile.cuh"
#include "writeback_pool.h"

namespace cebag {

namespace {

constexpr int kThreads = 256;

// per-call counters in the workspace (int32 each)
enum : int { kCtrUniqueHits = 0, kCtrMissLookups = 1, kCtrBadIndex = 2, kCtrUniqueMisses = 3, kCtrFlushed = 4,
             kCtrEvictable = 5,
             kCtrEvict = 6,      // E of this call (0 when the call is rejected)
             kCtrStatus = 7,     // device-side verdict: CEBAG_OK / CEBAG_ERR_CAPACITY / CEBAG_ERR_INDEX
             kCtrEpoch = 8,      // window stamp of this call = dev_state[EPOCH] + 1
             kCtrAdmit = 9,      // M of this call (0 when rejected)
             kCtrVictims = 10,   // victims ranked by row (== E)
             kCtrFixups = 11,    // positions to fix up (0 when rejected)
             kNumCounters = 16 };

struct SelectState {             // radix-select of the k smallest keys
    unsigned long long prefix;   // high digits of the k-th smallest key found so far
    long long k;                 // how many of the keys that match `prefix` are still to be taken
    int hist[256];
};

__device__ __forceinline__ int32_t row_of_id(const cebag_table& t, int64_t id) {
    return t.idx_map ? __ldg(t.idx_map + id) : (int32_t)id;
}

// ---- call bracket -------------------------------------------------------------------------------------------------
__global__ void begin_call_kernel(const cebag_table t, int32_t* __restrict__ counters) {
    if (threadIdx.x < kNumCounters) counters[threadIdx.x] = 0;
    __syncwarp();
    if (threadIdx.x == 0) counters[kCtrEpoch] = (int32_t)t.dev_state[CEBAG_STATE_EPOCH] + 1;
}

// The verdict of the call (A.3 assert, id range) and E, from the device-resident counts.  One CTA of 256 threads:
// thread 0 decides, everybody clears the radix-select histogram.
__global__ void __launch_bounds__(256)
decide_kernel(const cebag_table t, int32_t* __restrict__ counters, SelectState* __restrict__ sel) {
    sel->hist[threadIdx.x] = 0;
    if (threadIdx.x != 0) return;
    const long long M = counters[kCtrUniqueMisses], H = counters[kCtrUniqueHits];
    const long long avail = t.dev_state[CEBAG_STATE_AVAIL];
    int status = CEBAG_OK;
    long long E = 0;
    if (counters[kCtrBadIndex]) status = CEBAG_ERR_INDEX;
    else if (H + M > (long long)t.cache_rows) status = CEBAG_ERR_CAPACITY;
    else {
        E = M > avail ? M - avail : 0;
        // rows of an earlier, still protected window cannot be victims: enough others must exist
        if (t.protect_windows > 1 && E > (long long)counters[kCtrEvictable]) status = CEBAG_ERR_CAPACITY;
    }
    const bool ok = status == CEBAG_OK;
    counters[kCtrStatus] = status;
    counters[kCtrEvict] = ok ? (int32_t)E : 0;
    counters[kCtrAdmit] = ok ? (int32_t)M : 0;
    counters[kCtrFixups] = ok ? counters[kCtrMissLookups] : 0;
    sel->prefix = 0ull;
    sel->k = ok ? E : 0;
}

// Last kernel of the map work: the device-resident table state moves on (only if the call was accepted) and the
// result record goes to pinned host memory, status last.
__global__ void end_call_kernel(const cebag_table t, const int32_t* __restrict__ counters, int64_t n,
                                cebag_prepare_result* __restrict__ result) {
    if (threadIdx.x != 0) return;
    const int status = counters[kCtrStatus];
    const long long M = counters[kCtrAdmit], E = counters[kCtrEvict];
    if (status == CEBAG_OK) {
        t.dev_state[CEBAG_STATE_AVAIL] += E - M;
        t.dev_state[CEBAG_STATE_EPOCH] = counters[kCtrEpoch];
        // no LFU counter can grow by more than the n lookups of the call: a cheap upper bound for the victim selection
        if (t.freq) t.dev_state[CEBAG_STATE_MAXFREQ] += n;
    }
    t.dev_state[CEBAG_STATE_CALLS] += 1;
    if (result) {
        result->unique_hits = counters[kCtrUniqueHits];
        result->unique_misses = counters[kCtrUniqueMisses];
        result->evicted = E;
        result->miss_lookups = counters[kCtrMissLookups];
        result->total_lookups = n;
        result->evictable = counters[kCtrEvictable];
        result->avail_after = t.dev_state[CEBAG_STATE_AVAIL];
        __threadfence_system();
        *reinterpret_cast<volatile int64_t*>(&result->status) = status;
        __threadfence_system();
    }
}

// a rejected call must leave the miss bitmap all-zero again (an accepted one clears its bits as it commits)
__global__ void __launch_bounds__(kThreads)
clear_bitmap_if_rejected_kernel(const cebag_table t, const int32_t* __restrict__ counters, int64_t words) {
    if (counters[kCtrStatus] == CEBAG_OK) return;
    for (int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x; w < words; w += (int64_t)gridDim.x * kThreads)
        t.miss_bitmap[w] = 0u;
}

// ---- probe ---------------------------------------------------------------------------------------------------------
constexpr int kProbeIds = 4;  // ids in flight per thread

// Where the probe raises the hit flag of slot s.  The warm-up hands out slots by frequency rank and LFU keeps the hot
// rows where they are, so the hot slots are the LOW slot numbers: with one flag byte per slot in slot order the flags
// of the ~10^4 hottest slots -- 40 % of a Criteo window's 13.6 M lookups -- live in a few KB, i.e. in a handful of L2
// slices, and those slices serialise the whole kernel (ncu: one slice served 11 x the average number of sectors, issue
// slots busy 3 %).  The probe therefore writes a SPREAD copy of the flags in the call's workspace -- slot s at byte
// (s mod L) * M + (s div L), M = ceil(C / L): neighbouring slots are M bytes apart and the hot set covers the whole
// array -- and collect_hits_kernel moves the raised flags into the table's slot-ordered hit_flags, which is what every
// later kernel streams through.  L = 1 is the identity layout.
struct FlagLayout {
    int32_t log_l;     // L = 1 << log_l
    int32_t m;         // M
    int64_t bytes;     // L * M, a multiple of 16
};
constexpr int kMaxFlagSpreadLog = 16;

FlagLayout flag_layout(int64_t cache_rows, int spread) {
    FlagLayout f;
    f.log_l = 0;
    while ((1 << (f.log_l + 1)) <= spread && f.log_l + 1 <= kMaxFlagSpreadLog) ++f.log_l;
    const int64_t l = (int64_t)1 << f.log_l;
    int64_t m = (cache_rows + l - 1) / l;
    if (l == 1) m = (m + 15) / 16 * 16;
    f.m = (int32_t)m;
    f.bytes = (l * m + 15) / 16 * 16;
    return f;
}
// room for any layout the knob can ask for
size_t flag_spread_capacity(int64_t cache_rows) { return (size_t)cache_rows + ((size_t)1 << kMaxFlagSpreadLog) + 16; }

__device__ __forceinline__ int64_t flag_index(int32_t slot, int log_l, int m) {
    return (int64_t)(slot & ((1 << log_l) - 1)) * m + (slot >> log_l);
}

// MODE 0: one lane per distinct slot of the warp (match.any) checks the flag and raises it if it reads 0
// MODE 1: every lane that hit checks its flag (the four loads of a thread are issued back to back), raises it on a 0
// The check may read a stale 0 from L1, which only costs a store; racing stores of the same 1 are fine.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
probe_kernel(const cebag_table t, const int64_t* __restrict__ ids, int64_t n, int64_t* __restrict__ out,
             int32_t* __restrict__ miss_pos, int32_t* __restrict__ counters, uint8_t* __restrict__ flags, int log_l, int m) {
    __shared__ int32_t cta_misses, cta_base;
    const int lane = lane_id();
    if (threadIdx.x == 0) cta_misses = 0;
    __syncthreads();
    const int64_t tile = (int64_t)kThreads * kProbeIds;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile) {
        int64_t id[kProbeIds];
        int32_t row[kProbeIds], slot[kProbeIds];
        bool live[kProbeIds];
#pragma unroll
        for (int u = 0; u < kProbeIds; ++u) {
            int64_t i = base + (int64_t)u * kThreads + threadIdx.x;
            live[u] = i < n;
            id[u] = live[u] ? __ldg(ids + i) : 0;
        }
        bool bad = false;
#pragma unroll
        for (int u = 0; u < kProbeIds; ++u) {
            if (live[u] && (id[u] < 0 || id[u] >= t.num_rows)) { bad = true; live[u] = false; id[u] = 0; }
            row[u] = live[u] ? row_of_id(t, id[u]) : 0;
        }
        if (bad) atomicOr(&counters[kCtrBadIndex], 1);
#pragma unroll
        for (int u = 0; u < kProbeIds; ++u) slot[u] = live[u] ? t.row2slot[row[u]] : 0;
        unsigned missed[kProbeIds];
        int warp_misses = 0;
        if (MODE == 1) {
            bool hit[kProbeIds];
            uint8_t* fp[kProbeIds];
            unsigned seen[kProbeIds];
#pragma unroll
            for (int u = 0; u < kProbeIds; ++u) {
                const int64_t i = base + (int64_t)u * kThreads + threadIdx.x;
                const bool miss = live[u] && slot[u] < 0;
                hit[u] = live[u] && !miss;
                if (hit[u]) out[i] = slot[u];
                fp[u] = flags + (hit[u] ? flag_index(slot[u], log_l, m) : 0);
                missed[u] = __ballot_sync(0xffffffffu, miss);
                warp_misses += __popc(missed[u]);
            }
#pragma unroll
            for (int u = 0; u < kProbeIds; ++u) seen[u] = hit[u] ? (unsigned)*fp[u] : 1u;
#pragma unroll
            for (int u = 0; u < kProbeIds; ++u)
                if (!seen[u]) *fp[u] = 1;
        } else {
#pragma unroll
            for (int u = 0; u < kProbeIds; ++u) {
                const int64_t i = base + (int64_t)u * kThreads + threadIdx.x;
                const bool miss = live[u] && slot[u] < 0;
                const bool hit = live[u] && !miss;
                if (hit) out[i] = slot[u];
                // ids arrive feature-major: the 32 ids of a warp belong to one table, and for the small tables most of
                // them are the same few rows
                const unsigned same = __match_any_sync(0xffffffffu, hit ? slot[u] : -1 - lane);
                if (hit && lane == __ffs(same) - 1) {
                    uint8_t* flag = flags + flag_index(slot[u], log_l, m);
                    if (!*flag) *flag = 1;
                }
                missed[u] = __ballot_sync(0xffffffffu, miss);
                warp_misses += __popc(missed[u]);
            }
        }
        // Append the positions that missed: warps reserve inside the CTA (shared atomic), the CTA reserves in the list
        // with ONE global atomic per tile -- a single hot counter otherwise serialises ~10^5 warp-level atomics.
        int32_t warp_off = 0;
        if (warp_misses && lane == 0) warp_off = atomicAdd(&cta_misses, warp_misses);
        __syncthreads();
        if (threadIdx.x == 0) {
            cta_base = cta_misses ? atomicAdd(&counters[kCtrMissLookups], cta_misses) : 0;
            cta_misses = 0;
        }
        __syncthreads();
        if (warp_misses) {
            int32_t pos = cta_base + __shfl_sync(0xffffffffu, warp_off, 0);
#pragma unroll
            for (int u = 0; u < kProbeIds; ++u) {
                if ((missed[u] >> lane) & 1u) {
                    const int64_t i = base + (int64_t)u * kThreads + threadIdx.x;
                    out[i] = -1;
                    miss_pos[pos + __popc(missed[u] & ((1u << lane) - 1u))] = (int32_t)i;
                    const uint32_t bit = 1u << (row[u] & 31);
                    uint32_t* word = t.miss_bitmap + (row[u] >> 5);
                    if (!(*reinterpret_cast<volatile uint32_t*>(word) & bit)) atomicOr(word, bit);
                }
                pos += __popc(missed[u]);
            }
        }
    }
}

// The raised flags of the spread array go to the table's slot-ordered hit_flags (all-zero before); their number is the
// call's count of unique hits.  16 flags per 128-bit load.
__global__ void __launch_bounds__(kThreads)
collect_hits_kernel(const cebag_table t, const uint8_t* __restrict__ flags, int64_t vecs, int log_l, int m,
                    int32_t* __restrict__ counters) {
    const uint4* f16 = reinterpret_cast<const uint4*>(flags);
    int32_t c = 0;
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < vecs; v += (int64_t)gridDim.x * kThreads) {
        const uint4 f = f16[v];
        if (!(f.x | f.y | f.z | f.w)) continue;
        const uint32_t w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if ((w[k] >> (8 * b)) & 0xffu) {
                    const int64_t idx = v * 16 + k * 4 + b;
                    const int64_t slot = ((idx % m) << log_l) | (idx / m);
                    t.hit_flags[slot] = 1;
                    ++c;
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane_id() == 0 && c) atomicAdd(&counters[kCtrUniqueHits], c);
}

// ---- bitmap ranking: rows in ascending order -----------------------------------------------------------------------
// Used twice per call: for the missed rows (gate = the count of ids that missed) and, after the commit, for the
// victims' host rows (gate = E).  A zero gate means an empty bitmap: nothing is read.
constexpr int kWordsPerThread = 16;
constexpr int kWordsPerBlock = kScanThreads * kWordsPerThread;   // 4096 words = 131072 rows per CTA

__global__ void __launch_bounds__(kScanThreads)
bitmap_count_kernel(const uint32_t* __restrict__ bitmap, int64_t words, int32_t* __restrict__ block_sums,
                    const int32_t* __restrict__ gate) {
    __shared__ int32_t warp_sums[32];
    if (*gate == 0) {
        if (threadIdx.x == 0) block_sums[blockIdx.x] = 0;
        return;
    }
    int64_t w0 = (int64_t)blockIdx.x * kWordsPerBlock + (int64_t)threadIdx.x * kWordsPerThread;
    int32_t c = 0;
#pragma unroll
    for (int k = 0; k < kWordsPerThread; ++k) c += (w0 + k < words) ? __popc(bitmap[w0 + k]) : 0;
    int32_t tot = block_reduce_sum<kScanThreads>(c, warp_sums);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
bitmap_emit_kernel(const uint32_t* __restrict__ bitmap, int64_t words, const int32_t* __restrict__ block_base,
                   int32_t* __restrict__ rows_out, int64_t capacity, const int32_t* __restrict__ gate) {
    __shared__ int32_t warp_sums[32];
    if (*gate == 0) return;
    int64_t w0 = (int64_t)blockIdx.x * kWordsPerBlock + (int64_t)threadIdx.x * kWordsPerThread;
    uint32_t w[kWordsPerThread];
    int32_t c = 0;
#pragma unroll
    for (int k = 0; k < kWordsPerThread; ++k) {
        w[k] = (w0 + k < words) ? bitmap[w0 + k] : 0u;
        c += __popc(w[k]);
    }
    int64_t pos = (int64_t)block_base[blockIdx.x] + block_exclusive_scan<kScanThreads>(c, warp_sums);
    if (c == 0) return;
#pragma unroll
    for (int k = 0; k < kWordsPerThread; ++k) {
        uint32_t bits = w[k];
        while (bits) {
            int b = __ffs(bits) - 1;
            bits &= bits - 1;
            if (pos < capacity) rows_out[pos] = (int32_t)((w0 + k) * 32 + b);
            ++pos;
        }
    }
}

// ---- victim selection ----------------------------------------------------------------------------------------------------
// key of slot s if it may be evicted by the call stamped `epoch`: occupied, not hit by this call, not stamped by a
// protected window
__device__ __forceinline__ bool slot_key(const cebag_table& t, int64_t s, int32_t epoch, unsigned long long* key) {
    int32_t row = t.slot2row[s];
    if (row < 0) return false;                                                  // empty
    if (t.hit_flags[s]) return false;                                            // needed by this call
    const int32_t stamp = t.slot_epoch[s];
    if (stamp != 0 && epoch - stamp < t.protect_windows) return false;          // needed by a protected window
    if (t.strategy == CEBAG_EVICT_LFU) *key = (unsigned long long)t.freq[s];    // smallest counter first
    else *key = (unsigned long long)(0xffffffffu - (uint32_t)row);              // largest row first
    return true;
}

// LFU: every counter of an occupied slot is <= dev_state[MAXFREQ] (largest warm-start count + all lookups since), so a
// radix pass above that bound's highest byte sees digit 0 everywhere and changes nothing (the counters are int64)
__device__ __forceinline__ bool select_pass_is_void(const cebag_table& t, int shift) {
    return t.strategy == CEBAG_EVICT_LFU && shift > 0 &&
           ((unsigned long long)t.dev_state[CEBAG_STATE_MAXFREQ] >> shift) == 0ull;
}

// how many occupied slots may be evicted (only needed when more than the current window is protected)
__global__ void __launch_bounds__(kThreads)
count_evictable_kernel(const cebag_table t, int32_t* __restrict__ counters) {
    if (counters[kCtrMissLookups] == 0) return;
    const int32_t epoch = counters[kCtrEpoch];
    int32_t c = 0;
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < t.cache_rows; s += (int64_t)gridDim.x * kThreads) {
        unsigned long long key;
        c += slot_key(t, s, epoch, &key) ? 1 : 0;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane_id() == 0 && c) atomicAdd(&counters[kCtrEvictable], c);
}

__global__ void __launch_bounds__(kThreads)
select_hist_kernel(const cebag_table t, const int32_t* __restrict__ counters, SelectState* __restrict__ st, int shift,
                   int first_pass) {
    __shared__ int h[256];
    if (counters[kCtrEvict] == 0) return;
    if (select_pass_is_void(t, shift)) return;
    const int32_t epoch = counters[kCtrEpoch];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix;
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < t.cache_rows; s += (int64_t)gridDim.x * kThreads) {
        unsigned long long key;
        if (!slot_key(t, s, epoch, &key)) continue;
        if (!first_pass && (key >> (shift + 8)) != (prefix >> (shift + 8))) continue;
        atomicAdd(&h[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
select_choose_kernel(const cebag_table t, const int32_t* __restrict__ counters, SelectState* __restrict__ st, int shift) {
    // 256 threads, one per bin: the digit whose cumulative count first reaches k
    __shared__ int32_t warp_sums[32];
    if (counters[kCtrEvict] == 0) return;
    if (select_pass_is_void(t, shift)) return;
    const long long k = st->k;
    const int32_t c = st->hist[threadIdx.x];
    const int32_t excl = block_exclusive_scan<256>(c, warp_sums);
    __syncthreads();
    st->hist[threadIdx.x] = 0;
    if ((long long)excl < k && k <= (long long)excl + c) {
        st->prefix |= ((unsigned long long)threadIdx.x) << shift;
        st->k = k - excl;
    }
}

// eq[s] = 1 where an eligible slot's key equals the threshold (LFU ties)
__global__ void __launch_bounds__(kThreads)
select_equal_flags_kernel(const cebag_table t, const int32_t* __restrict__ counters, const SelectState* __restrict__ st,
                          int32_t* __restrict__ eq) {
    if (counters[kCtrEvict] == 0) return;
    const int32_t epoch = counters[kCtrEpoch];
    const unsigned long long thr = st->prefix;
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < t.cache_rows; s += (int64_t)gridDim.x * kThreads) {
        unsigned long long key;
        eq[s] = (slot_key(t, s, epoch, &key) && key == thr) ? 1 : 0;
    }
}

// free[s] = 1 for empty slots and for this call's victims
__global__ void __launch_bounds__(kThreads)
free_flags_kernel(const cebag_table t, const int32_t* __restrict__ counters, const SelectState* __restrict__ st,
                  const int32_t* __restrict__ eq_rank, int32_t* __restrict__ free_flag) {
    if (counters[kCtrAdmit] == 0) return;
    const bool evicting = counters[kCtrEvict] > 0;
    const int32_t epoch = counters[kCtrEpoch];
    unsigned long long thr = 0;
    long long take = 0;
    if (evicting) { thr = st->prefix; take = st->k; }
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < t.cache_rows; s += (int64_t)gridDim.x * kThreads) {
        int f = t.slot2row[s] < 0;
        if (!f && evicting) {
            unsigned long long key;
            if (slot_key(t, s, epoch, &key)) {
                if (key < thr) f = 1;
                else if (key == thr) f = eq_rank ? (eq_rank[s] < take) : 1;
            }
        }
        free_flag[s] = f;
    }
}

__global__ void __launch_bounds__(kThreads)
emit_free_slots_kernel(const int32_t* __restrict__ counters, const int32_t* __restrict__ free_flag_in,
                       const int32_t* __restrict__ free_pos, int64_t cache_rows, int32_t* __restrict__ free_slots) {
    const int32_t want = counters[kCtrAdmit];
    if (want == 0) return;
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < cache_rows; s += (int64_t)gridDim.x * kThreads) {
        if (free_flag_in[s] && free_pos[s] < want) free_slots[free_pos[s]] = (int32_t)s;
    }
}

// ---- commit ----------------------------------------------------------------------------------------------------------------
// Admission, maps: the j-th smallest missed row goes to the j-th lowest free slot (A.4 steps 4-6).  Records the row the
// slot held (the victim, or -1), updates both maps, the LFU counter, the window stamp and the miss bitmap.  The
// victim's row2slot entry becomes the marker -2 - slot ("not resident; its last value is still in that slot") until
// resolve_victims has ranked the victims by host row.
__global__ void __launch_bounds__(kThreads)
commit_admission_kernel(const cebag_table t, const int32_t* __restrict__ counters,
                        const int32_t* __restrict__ miss_rows, const int32_t* __restrict__ free_slots,
                        int32_t* __restrict__ victim_rows, int32_t* __restrict__ fill_src) {
    const int64_t m = counters[kCtrAdmit];
    const int32_t epoch = counters[kCtrEpoch];
    for (int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x; j < m; j += (int64_t)gridDim.x * kThreads) {
        const int32_t row = miss_rows[j];
        const int32_t slot = free_slots[j];
        const int32_t old_row = t.slot2row[slot];
        victim_rows[j] = old_row;
        if (old_row >= 0) t.row2slot[old_row] = -2 - slot;
        // a marker left by an earlier call: the row's write-back may still be in flight, its last value is entry
        // -2 - marker of that call's staging buffer
        const int32_t before = t.row2slot[row];
        fill_src[j] = before <= -2 ? -2 - before : -1;
        t.slot2row[slot] = row;
        t.row2slot[row] = slot;
        t.slot_epoch[slot] = epoch;
        if (t.freq) t.freq[slot] = 0;
        t.miss_bitmap[row >> 5] = 0u;   // every row of this word was missed in this call and is being admitted
    }
}

// slots hit by this call get the window stamp (only if the call was accepted); the hit flags are all-zero again
__global__ void __launch_bounds__(kThreads)
stamp_hits_kernel(const cebag_table t, const int32_t* __restrict__ counters, int64_t words) {
    const bool ok = counters[kCtrStatus] == CEBAG_OK;
    const int32_t epoch = counters[kCtrEpoch];
    uint32_t* flags4 = reinterpret_cast<uint32_t*>(t.hit_flags);          // 4 slots per word
    for (int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x; w < words; w += (int64_t)gridDim.x * kThreads) {
        const uint32_t f = flags4[w];
        if (!f) continue;
        flags4[w] = 0u;
        if (!ok) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if ((f >> (8 * b)) & 0xffu) t.slot_epoch[w * 4 + b] = epoch;
    }
}

// the victims' host rows go into the (all-zero again) row bitmap; ranking it gives them in ascending row order
__global__ void __launch_bounds__(kThreads)
mark_victims_kernel(const cebag_table t, const int32_t* __restrict__ counters, const int32_t* __restrict__ victim_rows) {
    if (counters[kCtrEvict] == 0) return;
    const int64_t m = counters[kCtrAdmit];
    for (int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x; j < m; j += (int64_t)gridDim.x * kThreads) {
        const int32_t old_row = victim_rows[j];
        if (old_row >= 0) atomicOr(t.miss_bitmap + (old_row >> 5), 1u << (old_row & 31));
    }
}

// i-th victim by host row: which slot still holds its last value?  Takes the marker out of row2slot and the bit out
// of the bitmap.
__global__ void __launch_bounds__(kThreads)
resolve_victims_kernel(const cebag_table t, const int32_t* __restrict__ counters,
                       const int32_t* __restrict__ victims_sorted, int32_t* __restrict__ victim_slots,
                       int64_t marked_rows) {
    const int64_t e = counters[kCtrEvict];
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < e; i += (int64_t)gridDim.x * kThreads) {
        const int32_t row = victims_sorted[i];
        victim_slots[i] = -2 - t.row2slot[row];
        // parked victims whose write-back is asynchronous keep a marker (their index in the staging buffer)
        t.row2slot[row] = i < marked_rows ? (int32_t)(-2 - i) : -1;
        t.miss_bitmap[row >> 5] = 0u;   // every bit of this word is a victim of this call
    }
}

// markers of an earlier call whose write-back has completed: the host table is current again
__global__ void __launch_bounds__(kThreads)
retire_markers_kernel(const cebag_table t, const int32_t* __restrict__ old_counters,
                      const int32_t* __restrict__ old_victims_sorted, int64_t bound) {
    int64_t e = old_counters[kCtrEvict];
    if (e > bound) e = bound;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < e; i += (int64_t)gridDim.x * kThreads) {
        const int32_t row = old_victims_sorted[i];
        if (t.row2slot[row] == (int32_t)(-2 - i)) t.row2slot[row] = -1;      // not re-admitted since
    }
}

// flush: no marker survives (every write-back has completed by then)
__global__ void __launch_bounds__(kThreads) normalize_markers_kernel(const cebag_table t) {
    for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < t.num_rows; r += (int64_t)gridDim.x * kThreads)
        if (t.row2slot[r] < -1) t.row2slot[r] = -1;
}

// ---- row movement -----------------------------------------------------------------------------------------------------------
// copy one row of `dim` floats with the 32 lanes of a warp (128-bit when possible)
template <bool VEC>
__device__ __forceinline__ void warp_copy_row(float* __restrict__ dst, const float* __restrict__ src, int dim, int lane) {
    if (VEC) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int c = lane; c < dim / 4; c += 32) d4[c] = s4[c];
    } else {
        for (int c = lane; c < dim; c += 32) dst[c] = src[c];
    }
}

// Row traffic of one call, one warp per row.  All lists are in ascending HOST ROW order.
//   kParkVictims : cache[victim_slots[i]] -> stage[i]                       i in [0, min(E, stage_rows))    HBM -> HBM
//   kWriteDirect : cache[victim_slots[i]] -> host_table[victims_sorted[i]]  i in [first, E)                 D2H
//   kWriteParked : stage[i]               -> host_table[victims_sorted[i]]  i in [dma_rows, min(E, stage_rows))  D2H
//                  (entries below dma_rows leave through the copy engine + host threads instead)
//   kFill        : host_table[miss_rows[j]] -> cache[free_slots[j]]         j in [0, M)                     H2D
//                  or prev_stage[fill_src[j]] -> cache[free_slots[j]] for rows whose write-back may be in flight
// The row-wise Adagrad state travels with the row.
enum : int { kParkVictims = 0, kWriteDirect = 1, kWriteParked = 2, kFill = 3 };

struct RowLists {
    const int32_t* miss_rows;
    const int32_t* free_slots;
    const int32_t* victims_sorted;
    const int32_t* victim_slots;
    float* stage;
    float* stage_state;
    int64_t stage_rows;
    int64_t dma_rows;
    const int32_t* fill_src;
    const float* prev_stage;
    const float* prev_stage_state;
};

template <bool VEC, int WHAT>
__global__ void __launch_bounds__(kThreads)
move_window_rows_kernel(const cebag_table t, const int32_t* __restrict__ counters, const RowLists L) {
    const int lane = lane_id();
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t num_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int dim = t.dim;
    const bool with_state = t.host_state && t.cache_state;
    const int64_t e = counters[kCtrEvict];
    const int64_t parked = L.stage ? (e < L.stage_rows ? e : L.stage_rows) : 0;
    int64_t lo = 0, hi = 0;
    if (WHAT == kParkVictims) hi = parked;
    else if (WHAT == kWriteParked) { lo = L.dma_rows < parked ? L.dma_rows : parked; hi = parked; }
    else if (WHAT == kWriteDirect) { lo = parked; hi = e; }
    else hi = counters[kCtrAdmit];
    for (int64_t j = lo + warp; j < hi; j += num_warps) {
        if (WHAT == kFill) {
            const int32_t slot = L.free_slots[j], row = L.miss_rows[j];
            const int32_t fwd = L.prev_stage ? L.fill_src[j] : -1;
            const float* src = fwd >= 0 ? L.prev_stage + (int64_t)fwd * dim : t.host_table + (int64_t)row * dim;
            warp_copy_row<VEC>(t.cache + (int64_t)slot * dim, src, dim, lane);
            if (lane == 0 && with_state)
                t.cache_state[slot] = (fwd >= 0 && L.prev_stage_state) ? L.prev_stage_state[fwd] : t.host_state[row];
        } else if (WHAT == kParkVictims) {
            const int32_t slot = L.victim_slots[j];
            warp_copy_row<VEC>(L.stage + j * dim, t.cache + (int64_t)slot * dim, dim, lane);
            if (lane == 0 && with_state && L.stage_state) L.stage_state[j] = t.cache_state[slot];
        } else if (WHAT == kWriteDirect) {
            const int32_t slot = L.victim_slots[j], row = L.victims_sorted[j];
            warp_copy_row<VEC>(t.host_table + (int64_t)row * dim, t.cache + (int64_t)slot * dim, dim, lane);
            if (lane == 0 && with_state) t.host_state[row] = t.cache_state[slot];
        } else {
            const int32_t row = L.victims_sorted[j];
            warp_copy_row<VEC>(t.host_table + (int64_t)row * dim, L.stage + j * dim, dim, lane);
            if (lane == 0 && with_state && L.stage_state) t.host_state[row] = L.stage_state[j];
        }
    }
}

// direction 0: host -> cache (admit / preload), 1: cache -> host (evict).  rows/slots given explicitly.
template <bool VEC>
__global__ void __launch_bounds__(kThreads)
move_rows_kernel(const cebag_table t, const int32_t* __restrict__ rows, const int32_t* __restrict__ slots,
                 const int64_t* __restrict__ freq_init, int64_t k, int direction) {
    const int lane = lane_id();
    const int64_t warp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int64_t num_warps = ((int64_t)gridDim.x * kThreads) >> 5;
    for (int64_t j = warp; j < k; j += num_warps) {
        const int32_t slot = slots ? slots[j] : (int32_t)j;
        const int32_t row = direction == 0 ? rows[j] : t.slot2row[slot];
        if (row < 0) continue;
        float* crow = t.cache + (int64_t)slot * t.dim;
        float* hrow = t.host_table + (int64_t)row * t.dim;
        if (direction == 0) warp_copy_row<VEC>(crow, hrow, t.dim, lane);
        else warp_copy_row<VEC>(hrow, crow, t.dim, lane);
        if (lane == 0) {
            if (direction == 0) {
                if (t.host_state && t.cache_state) t.cache_state[slot] = t.host_state[row];
                t.slot2row[slot] = row;
                t.row2slot[row] = slot;
                if (t.freq) {
                    t.freq[slot] = freq_init ? freq_init[j] : 0;
                    if (freq_init && freq_init[j] > t.dev_state[CEBAG_STATE_MAXFREQ])
                        atomicMax(reinterpret_cast<unsigned long long*>(t.dev_state + CEBAG_STATE_MAXFREQ),
                                  (unsigned long long)freq_init[j]);
                }
            } else {
                if (t.host_state && t.cache_state) t.host_state[row] = t.cache_state[slot];
                t.slot2row[slot] = -1;
                t.row2slot[row] = -1;
                if (t.freq) t.freq[slot] = CEBAG_FREQ_EMPTY;
            }
        }
    }
}

__global__ void add_avail_kernel(const cebag_table t, long long delta) {
    if (threadIdx.x == 0 && blockIdx.x == 0) t.dev_state[CEBAG_STATE_AVAIL] += delta;
}

// flush: every resident row goes back to the host table, maps are emptied
template <bool VEC>
__global__ void __launch_bounds__(kThreads)
flush_kernel(const cebag_table t, int32_t* __restrict__ counters) {
    const int lane = lane_id();
    const int64_t warp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int64_t num_warps = ((int64_t)gridDim.x * kThreads) >> 5;
    int32_t moved = 0;
    for (int64_t s = warp; s < t.cache_rows; s += num_warps) {
        const int32_t row = t.slot2row[s];
        if (row >= 0) {
            warp_copy_row<VEC>(t.host_table + (int64_t)row * t.dim, t.cache + s * t.dim, t.dim, lane);
            if (lane == 0) {
                if (t.host_state && t.cache_state) t.host_state[row] = t.cache_state[s];
                t.row2slot[row] = -1;
                t.slot2row[s] = -1;
                ++moved;
            }
        }
        if (lane == 0) {
            if (t.freq) t.freq[s] = CEBAG_FREQ_EMPTY;
            t.slot_epoch[s] = 0;
        }
    }
    if (lane == 0 && moved) atomicAdd(&counters[kCtrFlushed], moved);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        t.dev_state[CEBAG_STATE_AVAIL] = t.cache_rows;
        t.dev_state[CEBAG_STATE_EPOCH] = 0;
        t.dev_state[CEBAG_STATE_MAXFREQ] = 0;
    }
}

// ---- fix-up and LFU count ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
fixup_kernel(const cebag_table t, const int32_t* __restrict__ counters, const int64_t* __restrict__ ids,
             const int32_t* __restrict__ miss_pos, int64_t* __restrict__ out) {
    const int64_t count = counters[kCtrFixups];
    for (int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x; k < count; k += (int64_t)gridDim.x * kThreads) {
        int32_t i = miss_pos[k];
        out[i] = t.row2slot[row_of_id(t, ids[i])];
    }
}

// LFU counters += multiplicity.  Ids arrive feature-major, so a tile of consecutive ids hits few distinct slots for
// small tables and mostly distinct ones for big tables: each CTA first aggregates its tile in a shared-memory hash
// table (one shared atomic per id, one per warp when all 32 lanes agree), then issues one 64-bit global reduction per
// distinct slot -- hot slots no longer serialise millions of global atomics.
constexpr int kLfuTile = 2048;            // ids per CTA iteration
constexpr int kLfuTable = 4096;           // hash entries (load factor <= 0.5)

__global__ void __launch_bounds__(kThreads)
lfu_count_kernel(const cebag_table t, const int32_t* __restrict__ counters, const int64_t* __restrict__ slots, int64_t n) {
    __shared__ int s_key[kLfuTable];
    __shared__ int s_cnt[kLfuTable];
    if (counters[kCtrStatus] != CEBAG_OK) return;
    const int lane = lane_id();
    for (int64_t base = (int64_t)blockIdx.x * kLfuTile; base < n; base += (int64_t)gridDim.x * kLfuTile) {
        for (int e = threadIdx.x; e < kLfuTable; e += kThreads) { s_key[e] = -1; s_cnt[e] = 0; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kLfuTile / kThreads; ++k) {
            const int64_t i = base + k * kThreads + threadIdx.x;
            const bool live = i < n;
            const int s = live ? (int)slots[i] : -1 - lane;
            int same = 0;
            __match_all_sync(0xffffffffu, s, &same);
            int add = 1;
            if (same) { add = 32; if (lane != 0) add = 0; }
            if (live && add) {
                unsigned h = ((unsigned)s * 2654435761u) >> 20;          // 12 bits
                while (true) {
                    int prev = atomicCAS(&s_key[h], -1, s);
                    if (prev == -1 || prev == s) { atomicAdd(&s_cnt[h], add); break; }
                    h = (h + 1) & (kLfuTable - 1);
                }
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < kLfuTable; e += kThreads) {
            if (s_cnt[e]) atomicAdd(reinterpret_cast<unsigned long long*>(t.freq + s_key[e]), (unsigned long long)s_cnt[e]);
        }
        __syncthreads();
    }
}

struct PrepLayout {
    size_t counters, select, miss_pos, miss_rows, free_slots, victim_rows, flags_a, flags_b, bitmap_sums, scan_ws, fill_src,
           spread_flags, total;
    int64_t bitmap_blocks, words, hit_words;
};

PrepLayout prep_layout(const cebag_table* t, int64_t n) {
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    PrepLayout L;
    int64_t nn = n > 0 ? n : 1;
    int64_t C = t->cache_rows;
    L.words = ceil_div(t->num_rows, 32);
    L.hit_words = ceil_div(C, 16) * 4;              // the hit flags as 32-bit words (4 slots each)
    L.bitmap_blocks = ceil_div(L.words, kWordsPerBlock);
    size_t off = 0;
    // everything but miss_pos sits at an offset that does not depend on n: a later call finds the counters and the
    // victim list of this one (retire_device) without knowing its n
    L.counters = off; off += align(kNumCounters * 4);
    L.select = off; off += align(sizeof(SelectState));
    L.miss_rows = off; off += align((size_t)C * 4);
    L.free_slots = off; off += align((size_t)C * 4);
    L.victim_rows = off; off += align((size_t)C * 4);
    L.flags_a = off; off += align((size_t)C * 4);      // free flags, later: the victims in ascending row order
    L.flags_b = off; off += align((size_t)C * 4);      // tie ranks / free positions, later: the victims' slots
    L.bitmap_sums = off; off += align((size_t)(L.bitmap_blocks + 1) * 4);
    int64_t longest = C > L.bitmap_blocks ? C : L.bitmap_blocks;
    L.scan_ws = off; off += align(scan_workspace_bytes(longest));
    L.fill_src = off; off += align((size_t)C * 4);
    L.spread_flags = off; off += align(flag_spread_capacity(C));
    L.miss_pos = off; off += align((size_t)nn * 4);
    L.total = off;
    return L;
}

int sgrid_for_retire(int64_t cache_rows) { return grid_for(cache_rows < 262144 ? cache_rows : 262144, kThreads, 8); }

bool table_vec_ok(const cebag_table* t) {
    return t->dim % 4 == 0 && aligned16(t->cache) && aligned16(t->host_table);
}

int check_table(const cebag_table* t) {
    CEBAG_REQUIRE(t != nullptr, "null table");
    CEBAG_REQUIRE(t->num_rows > 0 && t->num_rows < ((int64_t)1 << 31), "num_rows must be in (0, 2^31)");
    CEBAG_REQUIRE(t->dim > 0 && t->cache_rows > 0, "dim / cache_rows");
    CEBAG_REQUIRE(t->strategy == CEBAG_EVICT_LFU || t->strategy == CEBAG_EVICT_DATASET, "strategy");
    CEBAG_REQUIRE(t->host_table && t->cache && t->row2slot && t->slot2row && t->slot_epoch && t->miss_bitmap &&
                  t->hit_flags && t->dev_state, "table pointers");
    CEBAG_REQUIRE(aligned16(t->hit_flags), "hit_flags alignment");
    CEBAG_REQUIRE(t->strategy != CEBAG_EVICT_LFU || t->freq != nullptr, "LFU needs freq");
    CEBAG_REQUIRE(t->protect_windows >= 1 && t->protect_windows <= 1024, "protect_windows");
    return CEBAG_OK;
}

int read_state(const cebag_table* t, int64_t* state, cudaStream_t stream) {
    CEBAG_CUDA_CHECK(cudaMemcpyAsync(state, t->dev_state, CEBAG_STATE_WORDS * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                     stream));
    CEBAG_CUDA_CHECK(cudaStreamSynchronize(stream));
    return CEBAG_OK;
}

// single-row helpers take their (row, slot) by value
__global__ void move_one_kernel(const cebag_table t, int32_t row, int32_t slot, int direction, int vec) {
    const int lane = lane_id();
    int32_t r = direction == 0 ? row : t.slot2row[slot];
    if (r < 0) return;
    float* crow = t.cache + (int64_t)slot * t.dim;
    float* hrow = t.host_table + (int64_t)r * t.dim;
    if (direction == 0) { if (vec) warp_copy_row<true>(crow, hrow, t.dim, lane); else warp_copy_row<false>(crow, hrow, t.dim, lane); }
    else { if (vec) warp_copy_row<true>(hrow, crow, t.dim, lane); else warp_copy_row<false>(hrow, crow, t.dim, lane); }
    if (lane == 0) {
        if (direction == 0) {
            if (t.host_state && t.cache_state) t.cache_state[slot] = t.host_state[r];
            t.slot2row[slot] = r; t.row2slot[r] = slot;
            if (t.freq) t.freq[slot] = 0;
            t.dev_state[CEBAG_STATE_AVAIL] -= 1;
        } else {
            if (t.host_state && t.cache_state) t.host_state[r] = t.cache_state[slot];
            t.slot2row[slot] = -1; t.row2slot[r] = -1;
            if (t.freq) t.freq[slot] = CEBAG_FREQ_EMPTY;
            t.dev_state[CEBAG_STATE_AVAIL] += 1;
        }
    }
}

template <int WHAT>
int launch_rows(const cebag_table* t, const int32_t* counters, const RowLists& lists, int grid, int threads,
                cudaStream_t stream) {
    if (table_vec_ok(t) && (WHAT == kWriteDirect || (WHAT == kFill && aligned16(lists.prev_stage)) ||
                            (WHAT != kFill && aligned16(lists.stage))))
        move_window_rows_kernel<true, WHAT><<<grid, threads, 0, stream>>>(*t, counters, lists);
    else
        move_window_rows_kernel<false, WHAT><<<grid, threads, 0, stream>>>(*t, counters, lists);
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}

}  // namespace
}  // namespace cebag

using namespace cebag;

extern "C" size_t cebag_prepare_workspace_bytes(const cebag_table* t, int64_t n_ids) {
    if (!t) return 0;
    return prep_layout(t, n_ids).total;
}

extern "C" int cebag_prepare_ids_async(const cebag_table* t, const int64_t* ids, int64_t n, int64_t* slot_ids_out,
                                       const cebag_workspace* ws, cebag_prepare_result* result, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_table(t);
    if (rc) return rc;
    CEBAG_REQUIRE(ws && ws->device && result, "workspace / result");
    CEBAG_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "n");
    memset(result, 0, sizeof(*result));
    result->total_lookups = n;
    if (n == 0) {
        if (ws->copy_stream && ws->copy_done_event)
            CEBAG_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ws->copy_done_event),
                                             reinterpret_cast<cudaStream_t>(ws->copy_stream)));
        if (ws->copy_stream && ws->writeback_done_event)
            CEBAG_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ws->writeback_done_event),
                                             reinterpret_cast<cudaStream_t>(ws->copy_stream)));
        return CEBAG_OK;
    }
    CEBAG_REQUIRE(ids && slot_ids_out, "ids / out");
    PrepLayout L = prep_layout(t, n);
    CEBAG_REQUIRE(ws->device_bytes >= L.total, "prepare workspace too small");
    cebag_prepare_result* result_dev = nullptr;
    CEBAG_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&result_dev), result, 0));
    result->status = CEBAG_PREPARE_PENDING;
    char* base = reinterpret_cast<char*>(ws->device);
    int32_t* counters = reinterpret_cast<int32_t*>(base + L.counters);
    SelectState* sel = reinterpret_cast<SelectState*>(base + L.select);
    int32_t* miss_pos = reinterpret_cast<int32_t*>(base + L.miss_pos);
    int32_t* miss_rows = reinterpret_cast<int32_t*>(base + L.miss_rows);
    int32_t* free_slots = reinterpret_cast<int32_t*>(base + L.free_slots);
    int32_t* victim_rows = reinterpret_cast<int32_t*>(base + L.victim_rows);
    int32_t* flags_a = reinterpret_cast<int32_t*>(base + L.flags_a);
    int32_t* flags_b = reinterpret_cast<int32_t*>(base + L.flags_b);
    int32_t* bitmap_sums = reinterpret_cast<int32_t*>(base + L.bitmap_sums);
    int32_t* scan_ws = reinterpret_cast<int32_t*>(base + L.scan_ws);
    int32_t* fill_src = reinterpret_cast<int32_t*>(base + L.fill_src);
    const int64_t C = t->cache_rows;
    const int64_t admit_bound = n < C ? n : C;          // M <= min(n, C) whenever the call is accepted
    cudaStream_t cstream = ws->copy_stream ? reinterpret_cast<cudaStream_t>(ws->copy_stream) : stream;
    const bool park = cstream != stream && ws->stage != nullptr && ws->stage_rows > 0;
    // DMA write-back: the first dma_rows parked victims leave through a copy engine and the host threads
    const bool dma = park && ws->dma_stream != nullptr && ws->dma_rows > 0 && ws->dma_ring && ws->dma_ring_rows &&
                     ws->host_table_hostptr;
    const int64_t dma_rows = dma ? (ws->dma_rows < ws->stage_rows ? ws->dma_rows : ws->stage_rows) : 0;
    CEBAG_REQUIRE(!(ws->dma_stream && ws->dma_rows > 0) || dma, "DMA write-back needs copy_stream, stage, dma_ring, "
                  "dma_ring_rows and host_table_hostptr");
    CEBAG_REQUIRE(!dma || !(t->host_state && t->cache_state) || (ws->dma_ring_state && ws->stage_state && ws->host_state_hostptr),
                  "DMA write-back of a table with row state needs stage_state, dma_ring_state and host_state_hostptr");

    if (ws->retire_device) {   // markers of an earlier call: its write-back has completed, the host table is current
        const char* old = reinterpret_cast<const char*>(ws->retire_device);
        if (ws->retire_wait_event)
            CEBAG_CUDA_CHECK(cudaStreamWaitEvent(stream, reinterpret_cast<cudaEvent_t>(ws->retire_wait_event), 0));
        retire_markers_kernel<<<sgrid_for_retire(C), kThreads, 0, stream>>>(
            *t, reinterpret_cast<const int32_t*>(old + L.counters), reinterpret_cast<const int32_t*>(old + L.flags_a), C);
        count_launches(1);
        CEBAG_LAUNCH_CHECK();
    }
    // CTAs per SM of the map kernels (read per call).  They are latency-bound and, under the look-ahead driver, run next
    // to the bandwidth-bound forward / backward of the previous window, whose SM slots they take: narrow grids cost
    // the side stream time it has and give the compute stream time it needs.
    const int prep_ctas = env_int("CEBAG_PREP_CTAS_PER_SM", 8);
    const int sgrid = grid_for(C, kThreads, prep_ctas);
    const bool lfu = t->strategy == CEBAG_EVICT_LFU;

    {
        KernelScope scope(kKernProbe, stream, 3);
        begin_call_kernel<<<1, 32, 0, stream>>>(*t, counters);
        // tuning knobs, read per call (a call is ~40 launches; lets one process compare settings)
        const int probe_mode = env_int("CEBAG_PROBE_MODE", 1);
        const int probe_ctas = env_int("CEBAG_PROBE_CTAS_PER_SM", prep_ctas);
        const FlagLayout fl = flag_layout(C, env_int("CEBAG_FLAG_SPREAD", 4096));
        uint8_t* spread = reinterpret_cast<uint8_t*>(base + L.spread_flags);
        CEBAG_CUDA_CHECK(cudaMemsetAsync(spread, 0, (size_t)fl.bytes, stream));
        const int pgrid = grid_for(ceil_div(n, kProbeIds), kThreads, probe_ctas);
        if (probe_mode == 1)
            probe_kernel<1><<<pgrid, kThreads, 0, stream>>>(*t, ids, n, slot_ids_out, miss_pos, counters, spread, fl.log_l, fl.m);
        else
            probe_kernel<0><<<pgrid, kThreads, 0, stream>>>(*t, ids, n, slot_ids_out, miss_pos, counters, spread, fl.log_l, fl.m);
        collect_hits_kernel<<<grid_for(fl.bytes / 16, kThreads, prep_ctas), kThreads, 0, stream>>>(*t, spread, fl.bytes / 16, fl.log_l,
                                                                                          fl.m, counters);
        CEBAG_LAUNCH_CHECK();
    }
    {   // missed rows, ascending: count, then (after the verdict) emit
        KernelScope scope(kKernBitmapRank, stream);
        bitmap_count_kernel<<<(int)L.bitmap_blocks, kScanThreads, 0, stream>>>(t->miss_bitmap, L.words, bitmap_sums,
                                                                               counters + kCtrMissLookups);
        CEBAG_LAUNCH_CHECK();
        rc = exclusive_scan_inplace(bitmap_sums, L.bitmap_blocks, counters + kCtrUniqueMisses, scan_ws, stream);
        if (rc) return rc;
    }
    {
        KernelScope scope(kKernSelect, stream, t->protect_windows > 1 ? 2 : 1);
        if (t->protect_windows > 1) count_evictable_kernel<<<sgrid, kThreads, 0, stream>>>(*t, counters);
        decide_kernel<<<1, 256, 0, stream>>>(*t, counters, sel);
        CEBAG_LAUNCH_CHECK();
    }
    {
        KernelScope scope(kKernBitmapRank, stream);
        bitmap_emit_kernel<<<(int)L.bitmap_blocks, kScanThreads, 0, stream>>>(t->miss_bitmap, L.words, bitmap_sums,
                                                                              miss_rows, C, counters + kCtrAdmit);
        CEBAG_LAUNCH_CHECK();
    }
    {   // victims: the E smallest keys (no-ops when E == 0)
        KernelScope scope(kKernSelect, stream, lfu ? 17 : 8);
        const int top = lfu ? 56 : 24;
        for (int shift = top; shift >= 0; shift -= 8) {
            select_hist_kernel<<<sgrid, kThreads, 0, stream>>>(*t, counters, sel, shift, shift == top);
            select_choose_kernel<<<1, 256, 0, stream>>>(*t, counters, sel, shift);
        }
        CEBAG_LAUNCH_CHECK();
        if (lfu) {   // ties at the threshold go to the lowest slots
            select_equal_flags_kernel<<<sgrid, kThreads, 0, stream>>>(*t, counters, sel, flags_b);
            CEBAG_LAUNCH_CHECK();
            rc = exclusive_scan_inplace(flags_b, C, nullptr, scan_ws, stream);
            if (rc) return rc;
        }
    }
    {
        KernelScope scope(kKernFreeSlots, stream, 4);
        free_flags_kernel<<<sgrid, kThreads, 0, stream>>>(*t, counters, sel, lfu ? flags_b : nullptr, flags_a);
        CEBAG_LAUNCH_CHECK();
        // flags_b is free again: positions = exclusive scan of the flags
        CEBAG_CUDA_CHECK(cudaMemcpyAsync(flags_b, flags_a, (size_t)C * 4, cudaMemcpyDeviceToDevice, stream));
        rc = exclusive_scan_inplace(flags_b, C, nullptr, scan_ws, stream);
        if (rc) return rc;
        emit_free_slots_kernel<<<sgrid, kThreads, 0, stream>>>(counters, flags_a, flags_b, C, free_slots);
        commit_admission_kernel<<<grid_for(admit_bound, kThreads, prep_ctas), kThreads, 0, stream>>>(*t, counters, miss_rows,
                                                                                            free_slots, victim_rows, fill_src);
        stamp_hits_kernel<<<grid_for(L.hit_words, kThreads, prep_ctas), kThreads, 0, stream>>>(*t, counters, L.hit_words);
        CEBAG_LAUNCH_CHECK();
    }
    {   // everything that only needs the maps: slot ids of the ids that missed, LFU counts, the result record
        KernelScope scope(kKernFixup, stream);
        fixup_kernel<<<grid_for(n, kThreads, prep_ctas), kThreads, 0, stream>>>(*t, counters, ids, miss_pos, slot_ids_out);
        CEBAG_LAUNCH_CHECK();
    }
    if (lfu) {
        KernelScope scope(kKernLfuCount, stream);
        lfu_count_kernel<<<grid_for(ceil_div(n, kLfuTile) * kThreads, kThreads, prep_ctas), kThreads, 0, stream>>>(
            *t, counters, slot_ids_out, n);
        CEBAG_LAUNCH_CHECK();
    }
    {   // victims in ascending host-row order (flags_a) and the slots that still hold them (flags_b)
        KernelScope scope(kKernVictimRank, stream, 6);
        mark_victims_kernel<<<grid_for(admit_bound, kThreads, prep_ctas), kThreads, 0, stream>>>(*t, counters, victim_rows);
        bitmap_count_kernel<<<(int)L.bitmap_blocks, kScanThreads, 0, stream>>>(t->miss_bitmap, L.words, bitmap_sums,
                                                                               counters + kCtrEvict);
        CEBAG_LAUNCH_CHECK();
        rc = exclusive_scan_inplace(bitmap_sums, L.bitmap_blocks, counters + kCtrVictims, scan_ws, stream);
        if (rc) return rc;
        bitmap_emit_kernel<<<(int)L.bitmap_blocks, kScanThreads, 0, stream>>>(t->miss_bitmap, L.words, bitmap_sums,
                                                                              flags_a, C, counters + kCtrEvict);
        resolve_victims_kernel<<<grid_for(admit_bound, kThreads, prep_ctas), kThreads, 0, stream>>>(*t, counters, flags_a, flags_b,
                                                                                            dma ? ws->stage_rows : 0);
        clear_bitmap_if_rejected_kernel<<<grid_for(L.words, kThreads, prep_ctas), kThreads, 0, stream>>>(*t, counters, L.words);
        end_call_kernel<<<1, 32, 0, stream>>>(*t, counters, n, result_dev);
        CEBAG_LAUNCH_CHECK();
    }

    // ---- rows ---------------------------------------------------------------------------------------------------------
    // PCIe-bound: ~50 GB/s x ~2 us of latency is ~110 KB in flight, a few hundred rows.  A SMALL grid matters: under
    // the look-ahead driver these kernels run next to the fwd/bwd kernels, and measured step time falls from 0.71 to
    // 0.60 ms going from 296 to 37 CTAs of 128 threads (a pure gather still reaches 46 of 51 GB/s).
    const int swap_ctas = env_int("CEBAG_SWAP_CTAS", 56);
    const int swap_threads = env_int("CEBAG_SWAP_THREADS", 128);
    RowLists lists;
    lists.miss_rows = miss_rows;
    lists.free_slots = free_slots;
    lists.victims_sorted = flags_a;
    lists.victim_slots = flags_b;
    lists.stage = park ? ws->stage : nullptr;
    lists.stage_state = park ? ws->stage_state : nullptr;
    lists.stage_rows = park ? ws->stage_rows : 0;
    lists.dma_rows = dma_rows;
    lists.fill_src = fill_src;
    lists.prev_stage = ws->prev_stage;
    lists.prev_stage_state = ws->prev_stage_state;
    // the victims' rows (and the slots the fill overwrites) may still be in use by an earlier window
    if (ws->victims_ready_event)
        CEBAG_CUDA_CHECK(cudaStreamWaitEvent(stream, reinterpret_cast<cudaEvent_t>(ws->victims_ready_event), 0));
    if (park) {
        KernelScope scope(kKernPark, stream);
        rc = launch_rows<kParkVictims>(t, counters, lists, kNumSMs * 4, kThreads, stream);   // HBM -> HBM
        if (rc) return rc;
    }
    if (cstream != stream) {
        // one event per host thread, re-recorded by every call: a wait captures the record that precedes it
        static thread_local cudaEvent_t committed = nullptr;
        if (!committed) CEBAG_CUDA_CHECK(cudaEventCreateWithFlags(&committed, cudaEventDisableTiming));
        CEBAG_CUDA_CHECK(cudaEventRecord(committed, stream));
        CEBAG_CUDA_CHECK(cudaStreamWaitEvent(cstream, committed, 0));
        if (dma) {
            cudaStream_t dstream = reinterpret_cast<cudaStream_t>(ws->dma_stream);
            CEBAG_CUDA_CHECK(cudaStreamWaitEvent(dstream, committed, 0));
            if (ws->dma_wait_event)
                CEBAG_CUDA_CHECK(cudaStreamWaitEvent(dstream, reinterpret_cast<cudaEvent_t>(ws->dma_wait_event), 0));
            const size_t row_bytes = (size_t)t->dim * sizeof(float);
            CEBAG_CUDA_CHECK(cudaMemcpyAsync(ws->dma_ring, ws->stage, (size_t)dma_rows * row_bytes, cudaMemcpyDeviceToHost, dstream));
            CEBAG_CUDA_CHECK(cudaMemcpyAsync(ws->dma_ring_rows, flags_a, (size_t)dma_rows * sizeof(int32_t),
                                             cudaMemcpyDeviceToHost, dstream));
            const bool with_state = t->host_state && t->cache_state;
            if (with_state)
                CEBAG_CUDA_CHECK(cudaMemcpyAsync(ws->dma_ring_state, ws->stage_state, (size_t)dma_rows * sizeof(float),
                                                 cudaMemcpyDeviceToHost, dstream));
            WritebackJob* job = new WritebackJob();
            job->host_table = ws->host_table_hostptr;
            job->host_state = with_state ? ws->host_state_hostptr : nullptr;
            job->ring = ws->dma_ring;
            job->ring_state = with_state ? ws->dma_ring_state : nullptr;
            job->rows = ws->dma_ring_rows;
            job->ring_rows = dma_rows;
            job->dim = t->dim;
            job->evicted = &result->evicted;
            job->status = &result->status;
            cudaError_t e = cudaLaunchHostFunc(dstream, [](void* p) {
                WritebackJob* j = reinterpret_cast<WritebackJob*>(p);
                run_writeback_job(*j);
                delete j;
            }, job);
            if (e != cudaSuccess) {
                delete job;
                CEBAG_CUDA_CHECK(e);
            }
            if (ws->dma_done_event)
                CEBAG_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ws->dma_done_event), dstream));
        }
    }
    {   // victims that are not parked must leave their slots before the fill overwrites them
        KernelScope scope(kKernWriteBack, cstream);
        rc = launch_rows<kWriteDirect>(t, counters, lists, swap_ctas, swap_threads, cstream);
        if (rc) return rc;
    }
    {
        KernelScope scope(kKernFillRows, cstream);
        rc = launch_rows<kFill>(t, counters, lists, swap_ctas, swap_threads, cstream);
        if (rc) return rc;
    }
    if (cstream != stream && ws->copy_done_event)
        CEBAG_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ws->copy_done_event), cstream));
    if (park) {
        KernelScope scope(kKernWriteBack, cstream);
        rc = launch_rows<kWriteParked>(t, counters, lists, swap_ctas, swap_threads, cstream);
        if (rc) return rc;
    }
    if (cstream != stream && ws->writeback_done_event)
        CEBAG_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ws->writeback_done_event), cstream));
    return CEBAG_OK;
}

extern "C" int cebag_prepare_result_status(const cebag_table* t, const cebag_prepare_result* r,
                                           cebag_prepare_stats* stats) {
    CEBAG_REQUIRE(t != nullptr && r != nullptr, "table / result");
    const int64_t status = *reinterpret_cast<const volatile int64_t*>(&r->status);
    if (status == CEBAG_PREPARE_PENDING) return CEBAG_PREPARE_PENDING;
    if (stats) {
        stats->unique_hits = r->unique_hits;
        stats->unique_misses = r->unique_misses;
        stats->evicted = r->evicted;
        stats->miss_lookups = r->miss_lookups;
        stats->total_lookups = r->total_lookups;
    }
    if (status == CEBAG_ERR_INDEX) {
        set_error("prepare_ids: an id is outside [0, %lld)", (long long)t->num_rows);
    } else if (status == CEBAG_ERR_CAPACITY) {
        const long long need = (long long)(r->unique_hits + r->unique_misses);
        if (need > (long long)t->cache_rows)
            set_error("You move %lld embedding rows from CPU to CUDA. It is larger than the capacity of the cache, "
                      "which at most contains %lld rows, Please increase cuda_row_num or decrease the training batch size.",
                      need, (long long)t->cache_rows);
        else
            set_error("You move %lld embedding rows from CPU to CUDA while %d look-ahead windows are protected: only "
                      "%lld of the %lld cached rows may be evicted. It is larger than the capacity of the cache, Please "
                      "increase cuda_row_num or decrease the training batch size.",
                      need, t->protect_windows, (long long)r->evictable, (long long)t->cache_rows);
    }
    return (int)status;
}

extern "C" int cebag_prepare_ids(const cebag_table* t, const int64_t* ids, int64_t n, int64_t* slot_ids_out,
                                 const cebag_workspace* ws, cebag_prepare_stats* stats, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(ws && ws->pinned && stats, "workspace / stats");
    cebag_prepare_result* result = reinterpret_cast<cebag_prepare_result*>(ws->pinned);
    int rc = cebag_prepare_ids_async(t, ids, n, slot_ids_out, ws, result, stream);
    if (rc) return rc;
    CEBAG_CUDA_CHECK(cudaStreamSynchronize(stream));
    return cebag_prepare_result_status(t, result, stats);
}

extern "C" int cebag_available_rows(const cebag_table* t, int64_t* avail_out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_table(t);
    if (rc) return rc;
    CEBAG_REQUIRE(avail_out != nullptr, "avail_out");
    int64_t state[CEBAG_STATE_WORDS];
    rc = read_state(t, state, stream);
    if (rc) return rc;
    *avail_out = state[CEBAG_STATE_AVAIL];
    return CEBAG_OK;
}

extern "C" int cebag_flush(const cebag_table* t, const cebag_workspace* ws, int64_t* rows_written, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_table(t);
    if (rc) return rc;
    CEBAG_REQUIRE(ws && ws->device && ws->pinned && ws->device_bytes >= kNumCounters * 4, "workspace");
    int32_t* counters = reinterpret_cast<int32_t*>(ws->device);
    CEBAG_CUDA_CHECK(cudaMemsetAsync(counters, 0, kNumCounters * 4, stream));
    const int grid = grid_for((int64_t)t->cache_rows * 32, kThreads, 8);
    {
        KernelScope scope(kKernFlush, stream);
        if (table_vec_ok(t)) flush_kernel<true><<<grid, kThreads, 0, stream>>>(*t, counters);
        else flush_kernel<false><<<grid, kThreads, 0, stream>>>(*t, counters);
        normalize_markers_kernel<<<grid_for(t->num_rows, kThreads, 8), kThreads, 0, stream>>>(*t);
        count_launches(1);
    }
    CEBAG_LAUNCH_CHECK();
    CEBAG_CUDA_CHECK(cudaMemcpyAsync(ws->pinned, counters, kNumCounters * 4, cudaMemcpyDeviceToHost, stream));
    CEBAG_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (rows_written) *rows_written = reinterpret_cast<int32_t*>(ws->pinned)[kCtrFlushed];
    return CEBAG_OK;
}

extern "C" int cebag_preload(const cebag_table* t, const int32_t* rows, const int64_t* freq_init, int64_t k, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_table(t);
    if (rc) return rc;
    int64_t state[CEBAG_STATE_WORDS];
    rc = read_state(t, state, stream);
    if (rc) return rc;
    CEBAG_REQUIRE(k >= 0 && k <= t->cache_rows, "preload count");
    CEBAG_REQUIRE(state[CEBAG_STATE_AVAIL] == t->cache_rows, "preload needs an empty cache");
    if (k == 0) return CEBAG_OK;
    CEBAG_REQUIRE(rows != nullptr, "rows");
    const int grid = grid_for(k * 32, kThreads, 8);
    {
        KernelScope scope(kKernMoveRows, stream, 2);
        if (table_vec_ok(t)) move_rows_kernel<true><<<grid, kThreads, 0, stream>>>(*t, rows, nullptr, freq_init, k, 0);
        else move_rows_kernel<false><<<grid, kThreads, 0, stream>>>(*t, rows, nullptr, freq_init, k, 0);
        add_avail_kernel<<<1, 1, 0, stream>>>(*t, -(long long)k);
    }
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}

extern "C" int cebag_admit_row(const cebag_table* t, int64_t row, int64_t slot, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_table(t);
    if (rc) return rc;
    CEBAG_REQUIRE(row >= 0 && row < t->num_rows && slot >= 0 && slot < t->cache_rows, "row / slot");
    int64_t state[CEBAG_STATE_WORDS];
    rc = read_state(t, state, stream);
    if (rc) return rc;
    CEBAG_REQUIRE(state[CEBAG_STATE_AVAIL] > 0, "no free slot");
    count_launches(1);
    move_one_kernel<<<1, 32, 0, stream>>>(*t, (int32_t)row, (int32_t)slot, 0, table_vec_ok(t) ? 1 : 0);
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}

extern "C" int cebag_evict_slot(const cebag_table* t, int64_t slot, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_table(t);
    if (rc) return rc;
    CEBAG_REQUIRE(slot >= 0 && slot < t->cache_rows, "slot");
    count_launches(1);
    move_one_kernel<<<1, 32, 0, stream>>>(*t, -1, (int32_t)slot, 1, table_vec_ok(t) ? 1 : 0);
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}
