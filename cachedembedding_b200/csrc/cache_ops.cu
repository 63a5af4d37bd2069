// The cache manager on the device: CachedParamMgr.prepare_ids / _prepare_rows_on_cuda / flush / reorder's preload
// (reference call site recsys/dlrm_main.py:259).
//
// The reference finds the unique rows of a window with sort-based unique/isin/topk over millions of ids and moves
// rows with CPU gather/scatter + staged memcpys.  Here:
//   * probe       : one pass over the ids.  Resident rows answer from the dense int32 row->slot map and stamp their
//                   slot with the window epoch (that stamp IS the evict backlist); missing rows set a bit in a per-row
//                   bitmap and append their position to a fix-up list.  No sort.
//   * rank        : popcount-scan of the bitmap emits the missed rows in ascending row order (the reference's
//                   sorted-unique contract) and their count M.
//   * victims     : radix-select of the E = M - free smallest (freq, slot) [LFU] / largest row [DATASET] keys among
//                   slots not stamped by this window; ties broken by ascending slot.
//   * swap        : i-th smallest missed row -> i-th lowest free slot.  One kernel moves both directions: the evicted
//                   row is stored straight into the pinned host table and the missed row is loaded straight from it
//                   (zero-copy over PCIe, 128-bit accesses), so both PCIe directions run concurrently and no CPU
//                   thread touches a row.
//   * fix-up      : ids that missed get their new slot; LFU counters += multiplicity.
// Integer / byte work throughout; HBM- (maps) and PCIe- (rows) bound.
#include "common.cuh"
#include "scan.cuh"
#include "profile.cuh"

namespace cebag {

namespace {

constexpr int kThreads = 256;

// counters read back by the host (int32 each)
enum : int { kCtrUniqueHits = 0, kCtrMissLookups = 1, kCtrBadIndex = 2, kCtrUniqueMisses = 3, kCtrFlushed = 4,
             kCtrEvictable = 5,
             kNumCounters = 16 };

struct SelectState {             // radix-select of the k smallest keys
    unsigned long long prefix;   // high digits of the k-th smallest key found so far
    long long k;                 // how many of the keys that match `prefix` are still to be taken
    int hist[256];
};

__device__ __forceinline__ int32_t row_of_id(const cebag_table& t, int64_t id) {
    return t.idx_map ? __ldg(t.idx_map + id) : (int32_t)id;
}

// ---- probe ---------------------------------------------------------------------------------------------------------
constexpr int kProbeIds = 4;  // ids in flight per thread

__global__ void __launch_bounds__(kThreads)
probe_kernel(const cebag_table t, const int64_t* __restrict__ ids, int64_t n, int64_t* __restrict__ out,
             int32_t* __restrict__ miss_pos, int32_t* __restrict__ counters) {
    const int lane = lane_id();
    int32_t uniq = 0;
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
#pragma unroll
        for (int u = 0; u < kProbeIds; ++u) {
            int64_t i = base + (int64_t)u * kThreads + threadIdx.x;
            bool miss = live[u] && slot[u] < 0;
            if (live[u] && !miss) {
                out[i] = slot[u];
                if (t.slot_epoch[slot[u]] != t.epoch) {
                    int32_t old = atomicExch(&t.slot_epoch[slot[u]], t.epoch);
                    uniq += (old != t.epoch);
                }
            }
            // warp-aggregated append of the positions that missed
            unsigned m = __ballot_sync(0xffffffffu, miss);
            if (m) {
                int32_t basepos = 0;
                if (lane == __ffs(m) - 1) basepos = atomicAdd(&counters[kCtrMissLookups], __popc(m));
                basepos = __shfl_sync(0xffffffffu, basepos, __ffs(m) - 1);
                if (miss) {
                    out[i] = -1;
                    miss_pos[basepos + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
                    uint32_t bit = 1u << (row[u] & 31);
                    uint32_t* word = t.miss_bitmap + (row[u] >> 5);
                    if (!(*reinterpret_cast<volatile uint32_t*>(word) & bit)) atomicOr(word, bit);
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) uniq += __shfl_xor_sync(0xffffffffu, uniq, d);
    if (lane == 0 && uniq) atomicAdd(&counters[kCtrUniqueHits], uniq);
}

// ---- bitmap ranking: missed rows in ascending order -----------------------------------------------------------------
constexpr int kWordsPerThread = 16;
constexpr int kWordsPerBlock = kScanThreads * kWordsPerThread;   // 4096 words = 131072 rows per CTA

__global__ void __launch_bounds__(kScanThreads)
bitmap_count_kernel(const uint32_t* __restrict__ bitmap, int64_t words, int32_t* __restrict__ block_sums) {
    __shared__ int32_t warp_sums[32];
    int64_t w0 = (int64_t)blockIdx.x * kWordsPerBlock + (int64_t)threadIdx.x * kWordsPerThread;
    int32_t c = 0;
#pragma unroll
    for (int k = 0; k < kWordsPerThread; ++k) c += (w0 + k < words) ? __popc(bitmap[w0 + k]) : 0;
    int32_t tot = block_reduce_sum<kScanThreads>(c, warp_sums);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
bitmap_emit_kernel(const uint32_t* __restrict__ bitmap, int64_t words, const int32_t* __restrict__ block_base,
                   int32_t* __restrict__ rows_out, int64_t capacity) {
    __shared__ int32_t warp_sums[32];
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
__device__ __forceinline__ bool slot_key(const cebag_table& t, int64_t s, unsigned long long* key) {
    int32_t row = t.slot2row[s];
    if (row < 0) return false;                                       // empty
    const int32_t stamp = t.slot_epoch[s];
    if (stamp != 0 && t.epoch - stamp < t.protect_windows) return false;   // needed by a protected window
    if (t.strategy == CEBAG_EVICT_LFU) *key = (unsigned long long)t.freq[s];   // smallest counter first
    else *key = (unsigned long long)(0xffffffffu - (uint32_t)row);             // largest row first
    return true;
}

// how many occupied slots may be evicted (only needed when more than the current window is protected)
__global__ void __launch_bounds__(kThreads)
count_evictable_kernel(const cebag_table t, int32_t* __restrict__ counters) {
    int32_t c = 0;
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < t.cache_rows; s += (int64_t)gridDim.x * kThreads) {
        unsigned long long key;
        c += slot_key(t, s, &key) ? 1 : 0;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane_id() == 0 && c) atomicAdd(&counters[kCtrEvictable], c);
}

__global__ void __launch_bounds__(kThreads)
select_hist_kernel(const cebag_table t, SelectState* __restrict__ st, int shift, int first_pass) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix;
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < t.cache_rows; s += (int64_t)gridDim.x * kThreads) {
        unsigned long long key;
        if (!slot_key(t, s, &key)) continue;
        if (!first_pass && (key >> (shift + 8)) != (prefix >> (shift + 8))) continue;
        atomicAdd(&h[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) select_choose_kernel(SelectState* __restrict__ st, int shift) {
    // 256 threads, one per bin: the digit whose cumulative count first reaches k
    __shared__ int32_t warp_sums[32];
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
select_equal_flags_kernel(const cebag_table t, const SelectState* __restrict__ st, int32_t* __restrict__ eq) {
    const unsigned long long thr = st->prefix;
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < t.cache_rows; s += (int64_t)gridDim.x * kThreads) {
        unsigned long long key;
        eq[s] = (slot_key(t, s, &key) && key == thr) ? 1 : 0;
    }
}

// free[s] = 1 for empty slots and for this call's victims
__global__ void __launch_bounds__(kThreads)
free_flags_kernel(const cebag_table t, const SelectState* __restrict__ st, const int32_t* __restrict__ eq_rank,
                  int evicting, int32_t* __restrict__ free_flag) {
    unsigned long long thr = 0;
    long long take = 0;
    if (evicting) { thr = st->prefix; take = st->k; }
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < t.cache_rows; s += (int64_t)gridDim.x * kThreads) {
        int f = t.slot2row[s] < 0;
        if (!f && evicting) {
            unsigned long long key;
            if (slot_key(t, s, &key)) {
                if (key < thr) f = 1;
                else if (key == thr) f = eq_rank ? (eq_rank[s] < take) : 1;
            }
        }
        free_flag[s] = f;
    }
}

__global__ void __launch_bounds__(kThreads)
emit_free_slots_kernel(const int32_t* __restrict__ free_flag_in, const int32_t* __restrict__ free_pos, int64_t cache_rows,
                       int32_t* __restrict__ free_slots, int64_t want) {
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < cache_rows; s += (int64_t)gridDim.x * kThreads) {
        if (free_flag_in[s] && free_pos[s] < want) free_slots[free_pos[s]] = (int32_t)s;
    }
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

// Admission, step 1 (maps): the j-th smallest missed row goes to the j-th lowest free slot (A.4 steps 4-6).  Records
// the row the slot held (the victim, or -1) for the copy kernels and updates both maps, the LFU counter, the window
// stamp and the miss bitmap.  Everything after this kernel that only needs the maps (fix-up of the missed positions,
// LFU counts, backward plans) can proceed while the rows are still moving.
__global__ void __launch_bounds__(kThreads)
commit_admission_kernel(const cebag_table t, const int32_t* __restrict__ miss_rows,
                        const int32_t* __restrict__ free_slots, int32_t* __restrict__ victim_rows, int64_t m) {
    for (int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x; j < m; j += (int64_t)gridDim.x * kThreads) {
        const int32_t row = miss_rows[j];
        const int32_t slot = free_slots[j];
        const int32_t old_row = t.slot2row[slot];
        victim_rows[j] = old_row;
        if (old_row >= 0) t.row2slot[old_row] = -1;
        t.slot2row[slot] = row;
        t.row2slot[row] = slot;
        t.slot_epoch[slot] = t.epoch;
        if (t.freq) t.freq[slot] = 0;
        t.miss_bitmap[row >> 5] = 0u;   // every row of this word was missed in this call and is being admitted
    }
}

// Admission, step 2 (rows), one warp per row with zero-copy 128-bit accesses to the pinned host table.
// direction 1: victims HBM -> host table (write-back), direction 0: missed rows host table -> HBM (fill).
// The two directions are separate launches: interleaving posted writes and reads from one kernel halves the PCIe
// throughput of both (measured: 20-24 GB/s each way fused vs 51 GB/s for a pure gather; PCIe reads may not pass writes).
template <bool VEC>
__global__ void __launch_bounds__(kThreads)
copy_rows_kernel(const cebag_table t, const int32_t* __restrict__ miss_rows, const int32_t* __restrict__ free_slots,
                 const int32_t* __restrict__ victim_rows, int64_t m, int direction) {
    const int lane = lane_id();
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t num_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int dim = t.dim;
    for (int64_t j = warp; j < m; j += num_warps) {
        const int32_t slot = free_slots[j];
        float* crow = t.cache + (int64_t)slot * dim;
        if (direction == 1) {
            const int32_t old_row = victim_rows[j];
            if (old_row < 0) continue;
            warp_copy_row<VEC>(t.host_table + (int64_t)old_row * dim, crow, dim, lane);
            if (lane == 0 && t.host_state && t.cache_state) t.host_state[old_row] = t.cache_state[slot];
        } else {
            const int32_t row = miss_rows[j];
            warp_copy_row<VEC>(crow, t.host_table + (int64_t)row * dim, dim, lane);
            if (lane == 0 && t.host_state && t.cache_state) t.cache_state[slot] = t.host_state[row];
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
                if (t.freq) t.freq[slot] = freq_init ? freq_init[j] : 0;
            } else {
                if (t.host_state && t.cache_state) t.host_state[row] = t.cache_state[slot];
                t.slot2row[slot] = -1;
                t.row2slot[row] = -1;
                if (t.freq) t.freq[slot] = CEBAG_FREQ_EMPTY;
            }
        }
    }
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
}

// ---- fix-up and LFU count ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
fixup_kernel(const cebag_table t, const int64_t* __restrict__ ids, const int32_t* __restrict__ miss_pos, int64_t count,
             int64_t* __restrict__ out) {
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
lfu_count_kernel(const cebag_table t, const int64_t* __restrict__ slots, int64_t n) {
    __shared__ int s_key[kLfuTable];
    __shared__ int s_cnt[kLfuTable];
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
    size_t counters, select, miss_pos, miss_rows, free_slots, victim_rows, flags_a, flags_b, bitmap_sums, scan_ws, total;
    int64_t bitmap_blocks, words;
};

PrepLayout prep_layout(const cebag_table* t, int64_t n) {
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    PrepLayout L;
    int64_t nn = n > 0 ? n : 1;
    int64_t C = t->cache_rows;
    L.words = ceil_div(t->num_rows, 32);
    L.bitmap_blocks = ceil_div(L.words, kWordsPerBlock);
    size_t off = 0;
    L.counters = off; off += align(kNumCounters * 4);
    L.select = off; off += align(sizeof(SelectState));
    L.miss_pos = off; off += align((size_t)nn * 4);
    L.miss_rows = off; off += align((size_t)C * 4);
    L.free_slots = off; off += align((size_t)C * 4);
    L.victim_rows = off; off += align((size_t)C * 4);
    L.flags_a = off; off += align((size_t)C * 4);
    L.flags_b = off; off += align((size_t)C * 4);
    L.bitmap_sums = off; off += align((size_t)(L.bitmap_blocks + 1) * 4);
    int64_t longest = C > L.bitmap_blocks ? C : L.bitmap_blocks;
    L.scan_ws = off; off += align(scan_workspace_bytes(longest));
    L.total = off;
    return L;
}

int read_counters(const int32_t* dev, int32_t* pinned, cudaStream_t stream) {
    CEBAG_CUDA_CHECK(cudaMemcpyAsync(pinned, dev, kNumCounters * 4, cudaMemcpyDeviceToHost, stream));
    CEBAG_CUDA_CHECK(cudaStreamSynchronize(stream));
    return CEBAG_OK;
}

bool table_vec_ok(const cebag_table* t) {
    return t->dim % 4 == 0 && aligned16(t->cache) && aligned16(t->host_table);
}

int check_table(const cebag_table* t) {
    CEBAG_REQUIRE(t != nullptr, "null table");
    CEBAG_REQUIRE(t->num_rows > 0 && t->num_rows < ((int64_t)1 << 31), "num_rows must be in (0, 2^31)");
    CEBAG_REQUIRE(t->dim > 0 && t->cache_rows > 0, "dim / cache_rows");
    CEBAG_REQUIRE(t->strategy == CEBAG_EVICT_LFU || t->strategy == CEBAG_EVICT_DATASET, "strategy");
    CEBAG_REQUIRE(t->host_table && t->cache && t->row2slot && t->slot2row && t->slot_epoch && t->miss_bitmap,
                  "table pointers");
    CEBAG_REQUIRE(t->strategy != CEBAG_EVICT_LFU || t->freq != nullptr, "LFU needs freq");
    CEBAG_REQUIRE(t->protect_windows >= 1 && t->protect_windows <= 1024, "protect_windows");
    return CEBAG_OK;
}

// single-row helpers take their (row, slot) by value: stage them in a tiny device buffer owned by the stream order
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
        } else {
            if (t.host_state && t.cache_state) t.host_state[r] = t.cache_state[slot];
            t.slot2row[slot] = -1; t.row2slot[r] = -1;
            if (t.freq) t.freq[slot] = CEBAG_FREQ_EMPTY;
        }
    }
}

}  // namespace
}  // namespace cebag

using namespace cebag;

extern "C" size_t cebag_prepare_workspace_bytes(const cebag_table* t, int64_t n_ids) {
    if (!t) return 0;
    return prep_layout(t, n_ids).total;
}

extern "C" int cebag_prepare_ids(cebag_table* t, const int64_t* ids, int64_t n, int64_t* slot_ids_out,
                                 const cebag_workspace* ws, cebag_prepare_stats* stats, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_table(t);
    if (rc) return rc;
    CEBAG_REQUIRE(ws && ws->device && ws->pinned && stats, "workspace / stats");
    CEBAG_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "n");
    stats->unique_hits = stats->unique_misses = stats->evicted = stats->miss_lookups = 0;
    stats->total_lookups = n;
    if (n == 0) return CEBAG_OK;
    CEBAG_REQUIRE(ids && slot_ids_out, "ids / out");
    PrepLayout L = prep_layout(t, n);
    CEBAG_REQUIRE(ws->device_bytes >= L.total, "prepare workspace too small");
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
    int32_t* host_ctr = reinterpret_cast<int32_t*>(ws->pinned);
    const int64_t C = t->cache_rows;

    // a new window: stamps of earlier windows stop protecting their slots
    t->epoch = t->epoch >= 0x7ffffff0 ? 1 : t->epoch + 1;
    if (t->epoch == 1) CEBAG_CUDA_CHECK(cudaMemsetAsync(t->slot_epoch, 0, (size_t)C * 4, stream));

    CEBAG_CUDA_CHECK(cudaMemsetAsync(counters, 0, kNumCounters * 4, stream));
    {
        KernelScope scope(kKernProbe, stream);
        probe_kernel<<<grid_for(ceil_div(n, kProbeIds), kThreads, 8), kThreads, 0, stream>>>(*t, ids, n, slot_ids_out,
                                                                                            miss_pos, counters);
    }
    CEBAG_LAUNCH_CHECK();
    rc = read_counters(counters, host_ctr, stream);
    if (rc) return rc;
    const int64_t miss_lookups = host_ctr[kCtrMissLookups];
    stats->unique_hits = host_ctr[kCtrUniqueHits];
    stats->miss_lookups = miss_lookups;
    auto clear_bitmap = [&]() { return cudaMemsetAsync(t->miss_bitmap, 0, (size_t)L.words * 4, stream); };
    if (host_ctr[kCtrBadIndex]) {
        clear_bitmap();
        set_error("prepare_ids: an id is outside [0, %lld)", (long long)t->num_rows);
        return CEBAG_ERR_INDEX;
    }
    if (miss_lookups > 0) {
        // missed rows, ascending
        {
            KernelScope scope(kKernBitmapRank, stream);
            bitmap_count_kernel<<<(int)L.bitmap_blocks, kScanThreads, 0, stream>>>(t->miss_bitmap, L.words, bitmap_sums);
            CEBAG_LAUNCH_CHECK();
            rc = exclusive_scan_inplace(bitmap_sums, L.bitmap_blocks, counters + kCtrUniqueMisses, scan_ws, stream);
            if (rc) return rc;
        }
        rc = read_counters(counters, host_ctr, stream);
        if (rc) return rc;
        const int64_t M = host_ctr[kCtrUniqueMisses];
        stats->unique_misses = M;
        if (stats->unique_hits + M > C) {   // A.3 capacity assert, before any state changes
            clear_bitmap();
            set_error("You move %lld embedding rows from CPU to CUDA. It is larger than the capacity of the cache, "
                      "which at most contains %lld rows, Please increase cuda_row_num or decrease the training batch size.",
                      (long long)(stats->unique_hits + M), (long long)C);
            return CEBAG_ERR_CAPACITY;
        }
        {
            KernelScope scope(kKernBitmapRank, stream);
            bitmap_emit_kernel<<<(int)L.bitmap_blocks, kScanThreads, 0, stream>>>(t->miss_bitmap, L.words, bitmap_sums,
                                                                                  miss_rows, C);
        }
        CEBAG_LAUNCH_CHECK();

        const int64_t E = M > t->avail ? M - t->avail : 0;
        const int sgrid = grid_for(C, kThreads, 8);
        if (E > 0 && t->protect_windows > 1) {
            // rows of an earlier, still protected window cannot be victims: make sure enough others exist
            count_evictable_kernel<<<sgrid, kThreads, 0, stream>>>(*t, counters);
            count_launches(1);
            CEBAG_LAUNCH_CHECK();
            rc = read_counters(counters, host_ctr, stream);
            if (rc) return rc;
            if (E > host_ctr[kCtrEvictable]) {
                clear_bitmap();
                set_error("You move %lld embedding rows from CPU to CUDA while %d look-ahead windows are protected: only "
                          "%lld of the %lld cached rows may be evicted but %lld are needed. It is larger than the capacity "
                          "of the cache, Please increase cuda_row_num or decrease the training batch size.",
                          (long long)(stats->unique_hits + M), t->protect_windows, (long long)host_ctr[kCtrEvictable],
                          (long long)C, (long long)E);
                return CEBAG_ERR_CAPACITY;
            }
        }
        const bool lfu = t->strategy == CEBAG_EVICT_LFU;
        if (E > 0) {
            KernelScope scope(kKernSelect, stream, lfu ? 17 : 8);
            SelectState init;
            memset(&init, 0, sizeof(init));
            init.k = E;
            CEBAG_CUDA_CHECK(cudaMemcpyAsync(sel, &init, sizeof(init), cudaMemcpyHostToDevice, stream));
            const int top = lfu ? 56 : 24;
            for (int shift = top; shift >= 0; shift -= 8) {
                select_hist_kernel<<<sgrid, kThreads, 0, stream>>>(*t, sel, shift, shift == top);
                select_choose_kernel<<<1, 256, 0, stream>>>(sel, shift);
            }
            CEBAG_LAUNCH_CHECK();
            if (lfu) {   // ties at the threshold go to the lowest slots
                select_equal_flags_kernel<<<sgrid, kThreads, 0, stream>>>(*t, sel, flags_b);
                CEBAG_LAUNCH_CHECK();
                rc = exclusive_scan_inplace(flags_b, C, nullptr, scan_ws, stream);
                if (rc) return rc;
            }
        }
        {
            KernelScope scope(kKernFreeSlots, stream, 2);
            free_flags_kernel<<<sgrid, kThreads, 0, stream>>>(*t, sel, (E > 0 && lfu) ? flags_b : nullptr,
                                                              E > 0 ? 1 : 0, flags_a);
            CEBAG_LAUNCH_CHECK();
            // flags_b is free again: positions = exclusive scan of the flags
            CEBAG_CUDA_CHECK(cudaMemcpyAsync(flags_b, flags_a, (size_t)C * 4, cudaMemcpyDeviceToDevice, stream));
            rc = exclusive_scan_inplace(flags_b, C, nullptr, scan_ws, stream);
            if (rc) return rc;
            emit_free_slots_kernel<<<sgrid, kThreads, 0, stream>>>(flags_a, flags_b, C, free_slots, M);
            CEBAG_LAUNCH_CHECK();
        }
        {
            KernelScope scope(kKernFreeSlots, stream);
            commit_admission_kernel<<<grid_for(M, kThreads, 8), kThreads, 0, stream>>>(*t, miss_rows, free_slots,
                                                                                     victim_rows, M);
        }
        CEBAG_LAUNCH_CHECK();
        {
            // PCIe-bound: ~56 GB/s x ~2 us of latency is ~110 KB in flight, a few hundred rows.  A SMALL grid matters:
            // under the look-ahead driver these kernels run next to the fwd/bwd kernels, and measured step time falls
            // from 0.71 to 0.60 ms going from 296 to 37 CTAs of 128 threads (a pure gather still reaches 47 of
            // 51 GB/s).  With a copy stream in the workspace the rows move there while `stream` goes on with the
            // map-only work.
            static const int swap_ctas = env_int("CEBAG_SWAP_CTAS", 56);
            static const int swap_threads = env_int("CEBAG_SWAP_THREADS", 128);
            cudaStream_t cstream = ws->copy_stream ? reinterpret_cast<cudaStream_t>(ws->copy_stream) : stream;
            if (cstream != stream) {
                cudaEvent_t committed;
                CEBAG_CUDA_CHECK(cudaEventCreateWithFlags(&committed, cudaEventDisableTiming));
                CEBAG_CUDA_CHECK(cudaEventRecord(committed, stream));
                CEBAG_CUDA_CHECK(cudaStreamWaitEvent(cstream, committed, 0));
                CEBAG_CUDA_CHECK(cudaEventDestroy(committed));   // released once the wait has been satisfied
            }
            const int64_t want = ceil_div(M * 32, swap_threads);
            const int mgrid = (int)(want < swap_ctas ? want : swap_ctas);
            const bool vec = table_vec_ok(t);
            {
                KernelScope scope(kKernSwapRows, cstream, E > 0 ? 2 : 1);
                if (E > 0) {
                    if (vec) copy_rows_kernel<true><<<mgrid, swap_threads, 0, cstream>>>(*t, miss_rows, free_slots, victim_rows, M, 1);
                    else copy_rows_kernel<false><<<mgrid, swap_threads, 0, cstream>>>(*t, miss_rows, free_slots, victim_rows, M, 1);
                }
                if (vec) copy_rows_kernel<true><<<mgrid, swap_threads, 0, cstream>>>(*t, miss_rows, free_slots, victim_rows, M, 0);
                else copy_rows_kernel<false><<<mgrid, swap_threads, 0, cstream>>>(*t, miss_rows, free_slots, victim_rows, M, 0);
            }
            if (cstream != stream && ws->copy_done_event)
                CEBAG_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ws->copy_done_event), cstream));
        }
        CEBAG_LAUNCH_CHECK();
        {
            KernelScope scope(kKernFixup, stream);
            fixup_kernel<<<grid_for(miss_lookups, kThreads, 8), kThreads, 0, stream>>>(*t, ids, miss_pos, miss_lookups,
                                                                                      slot_ids_out);
        }
        CEBAG_LAUNCH_CHECK();
        stats->evicted = E;
        t->avail += E - M;
    }
    if (t->strategy == CEBAG_EVICT_LFU) {
        KernelScope scope(kKernLfuCount, stream);
        lfu_count_kernel<<<grid_for(ceil_div(n, kLfuTile) * kThreads, kThreads, 8), kThreads, 0, stream>>>(*t, slot_ids_out, n);
        CEBAG_LAUNCH_CHECK();
    }
    return CEBAG_OK;
}

extern "C" int cebag_flush(cebag_table* t, const cebag_workspace* ws, int64_t* rows_written, void* stream_) {
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
    }
    CEBAG_LAUNCH_CHECK();
    rc = read_counters(counters, reinterpret_cast<int32_t*>(ws->pinned), stream);
    if (rc) return rc;
    if (rows_written) *rows_written = reinterpret_cast<int32_t*>(ws->pinned)[kCtrFlushed];
    t->avail = t->cache_rows;
    t->epoch = 0;
    return CEBAG_OK;
}

extern "C" int cebag_preload(cebag_table* t, const int32_t* rows, const int64_t* freq_init, int64_t k, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_table(t);
    if (rc) return rc;
    CEBAG_REQUIRE(k >= 0 && k <= t->cache_rows && k <= t->avail, "preload count");
    CEBAG_REQUIRE(t->avail == t->cache_rows, "preload needs an empty cache");
    if (k == 0) return CEBAG_OK;
    CEBAG_REQUIRE(rows != nullptr, "rows");
    const int grid = grid_for(k * 32, kThreads, 8);
    {
        KernelScope scope(kKernMoveRows, stream);
        if (table_vec_ok(t)) move_rows_kernel<true><<<grid, kThreads, 0, stream>>>(*t, rows, nullptr, freq_init, k, 0);
        else move_rows_kernel<false><<<grid, kThreads, 0, stream>>>(*t, rows, nullptr, freq_init, k, 0);
    }
    CEBAG_LAUNCH_CHECK();
    t->avail -= k;
    return CEBAG_OK;
}

extern "C" int cebag_admit_row(cebag_table* t, int64_t row, int64_t slot, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_table(t);
    if (rc) return rc;
    CEBAG_REQUIRE(row >= 0 && row < t->num_rows && slot >= 0 && slot < t->cache_rows, "row / slot");
    CEBAG_REQUIRE(t->avail > 0, "no free slot");
    count_launches(1);
    move_one_kernel<<<1, 32, 0, stream>>>(*t, (int32_t)row, (int32_t)slot, 0, table_vec_ok(t) ? 1 : 0);
    CEBAG_LAUNCH_CHECK();
    t->avail -= 1;
    return CEBAG_OK;
}

extern "C" int cebag_evict_slot(cebag_table* t, int64_t slot, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_table(t);
    if (rc) return rc;
    CEBAG_REQUIRE(slot >= 0 && slot < t->cache_rows, "slot");
    count_launches(1);
    move_one_kernel<<<1, 32, 0, stream>>>(*t, -1, (int32_t)slot, 1, table_vec_ok(t) ? 1 : 0);
    CEBAG_LAUNCH_CHECK();
    t->avail += 1;
    return CEBAG_OK;
}
