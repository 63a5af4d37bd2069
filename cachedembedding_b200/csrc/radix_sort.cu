// Stable LSD radix sort of (slot, value) pairs -- the integer preprocessing of the fused backward.
// Keys are slot ids (< C, so only ceil(log2 (C + 1)) bits are sorted); values are bag ids or lookup positions.
//
// One-sweep formulation: ONE histogram kernel counts the digits of every pass (keys are read once), one tiny kernel
// turns the counts into global digit offsets, and each pass is a single kernel -- a tile ranks its keys per digit
// (warp-level match.any, per-warp digit counters in shared memory), obtains the number of equal digits in all earlier
// tiles by DECOUPLED LOOK-BACK over a per-tile status array (tiles take their index from a ticket counter, so every
// tile a tile waits for is already running), and scatters.  8-bit digits: 256 digits = one look-back lane per thread.
// Per pass 8 B read + 8 B written per pair; 21-bit slot ids sort in 3 passes (+ 8 B/key for the histogram).
// Replaces round 1's histogram / 3-kernel scan / scatter per pass (8 launches, 121 us for 1.7 M pairs).
#include "common.cuh"
#include "radix_sort.cuh"
#include "profile.cuh"

namespace cebag {

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kHistItems = 16;                               // keys per thread and round of the histogram kernel
constexpr int kHistTile = kSortThreads * kHistItems;
constexpr int kDigitBits = 8;
constexpr int kDigits = 1 << kDigitBits;                     // == kSortThreads: thread b owns digit b
constexpr int kMaxPasses = 4;
static_assert(kDigits == kSortThreads, "one look-back lane per digit");

// look-back status word: 2 flag bits + 30 value bits
constexpr uint32_t kFlagAggregate = 1u << 30;                // value = this tile's count of the digit
constexpr uint32_t kFlagPrefix = 2u << 30;                   // value = count of the digit in this and all earlier tiles
constexpr uint32_t kValueMask = (1u << 30) - 1u;

template <bool FIRST>
__device__ __forceinline__ uint32_t load_key(const void* keys_in, int64_t i) {
    if (FIRST) return (uint32_t) reinterpret_cast<const int64_t*>(keys_in)[i];
    return reinterpret_cast<const uint32_t*>(keys_in)[i];
}

// digit counts of all passes in one read of the keys: hist[p * 256 + d]
template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
sort_histogram_kernel(const KeyT* __restrict__ slot_ids, int64_t n, int passes, int key_bits, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[kMaxPasses][kDigits];
    for (int p = 0; p < passes; ++p) h[p][threadIdx.x] = 0;
    __syncthreads();
    for (int64_t base = (int64_t)blockIdx.x * kHistTile; base < n; base += (int64_t)gridDim.x * kHistTile) {
#pragma unroll 4
        for (int k = 0; k < kHistItems; ++k) {
            const int64_t i = base + (int64_t)k * kSortThreads + threadIdx.x;
            const bool valid = i < n;
            const uint32_t key = valid ? (uint32_t)slot_ids[i] : 0u;
            for (int p = 0; p < passes; ++p) {
                const int shift = p * kDigitBits;
                const int bits = key_bits - shift < kDigitBits ? key_bits - shift : kDigitBits;
                const uint32_t d = (key >> shift) & ((1u << bits) - 1u);
                // a warp whose 32 keys share the digit (hot slots of small tables) adds once
                int same = 0;
                __match_all_sync(0xffffffffu, valid ? d : (0x80000000u | (uint32_t)lane_id()), &same);
                if (same) { if (lane_id() == 0 && valid) atomicAdd(&h[p][d], 32u); }
                else if (valid) atomicAdd(&h[p][d], 1u);
            }
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p)
        if (h[p][threadIdx.x]) atomicAdd(&hist[p * kDigits + threadIdx.x], h[p][threadIdx.x]);
}

// per pass: exclusive scan of the 256 digit counts -> first output position of every digit
__global__ void __launch_bounds__(kSortThreads) sort_offsets_kernel(uint32_t* __restrict__ hist, int passes) {
    __shared__ uint32_t warp_tot[kSortWarps];
    for (int p = 0; p < passes; ++p) {
        const uint32_t c = hist[p * kDigits + threadIdx.x];
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane_id() >= d) incl += t;
        }
        if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) base += warp_tot[w];
        hist[p * kDigits + threadIdx.x] = base + incl - c;
        __syncthreads();
    }
}

// One pass.  Tile order = warp-major, then round-major, then lane: position within the tile is
// warp * (32 * ITEMS) + round * 32 + lane, which the ranking preserves inside every digit (stable).
template <bool FIRST, int kSortItems>
__global__ void __launch_bounds__(kSortThreads, kSortItems <= 8 ? 6 : 3)
sort_onesweep_kernel(const void* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, int64_t n, int shift, int bits,
                     const uint32_t* __restrict__ digit_offset, uint32_t* __restrict__ tile_state,
                     uint32_t* __restrict__ ticket, uint32_t num_tiles, uint32_t* __restrict__ keys_out,
                     uint32_t* __restrict__ vals_out) {
    constexpr int kSortTile = kSortThreads * kSortItems;
    __shared__ uint16_t cnt[kSortWarps][kDigits];   // per-warp digit counts (<= 32 * kSortItems)
    __shared__ uint32_t tile_base[kDigits];         // output position of the tile's first key of every digit
    __shared__ uint32_t tile_s;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t mask = (1u << bits) - 1u;
    // A CTA takes tiles from the ticket counter until none is left (grid <= tiles: the launch may cap the number of
    // CTAs so that the sort leaves SM slots to the kernels it runs next to).  Tickets are handed out in order to CTAs
    // that are running, so every tile a look-back waits for is being worked on.
    while (true) {
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = tile_s;
    if (tile >= num_tiles) break;
    const int64_t wbase = (int64_t)tile * kSortTile + (int64_t)warp * (32 * kSortItems);
    uint32_t key[kSortItems], val[kSortItems];
    uint16_t rank[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        const bool valid = i < n;
        key[k] = valid ? load_key<FIRST>(keys_in, i) : 0u;
        val[k] = valid ? (vals_in ? vals_in[i] : (uint32_t)i) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const bool valid = wbase + k * 32 + lane < n;
        const uint32_t d = valid ? ((key[k] >> shift) & mask) : (0x80000000u | (uint32_t)lane);  // invalid: match only self
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = cnt[warp][d];
            cnt[warp][d] = (uint16_t)(old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = (uint16_t)(old + __popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    __syncthreads();
    {   // thread b: digit b.  Counts of the warps -> exclusive prefix over the warps + the tile's total
        const int b = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = cnt[w][b];
            cnt[w][b] = (uint16_t)run;
            run += c;
        }
        // decoupled look-back: how many keys with digit b do the earlier tiles hold?
        volatile uint32_t* state = tile_state + b;
        uint32_t excl = 0;
        if (tile == 0) {
            state[0] = kFlagPrefix | run;
        } else {
            state[(size_t)tile * kDigits] = kFlagAggregate | run;
            for (int64_t p = (int64_t)tile - 1; p >= 0; --p) {
                uint32_t v;
                do { v = state[(size_t)p * kDigits]; } while ((v & ~kValueMask) == 0u);
                excl += v & kValueMask;
                if (v & kFlagPrefix) break;
            }
            state[(size_t)tile * kDigits] = kFlagPrefix | (excl + run);
        }
        tile_base[b] = digit_offset[b] + excl;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        if (wbase + k * 32 + lane < n) {
            const uint32_t d = (key[k] >> shift) & mask;
            const uint32_t pos = tile_base[d] + cnt[warp][d] + rank[k];
            keys_out[pos] = key[k];
            vals_out[pos] = val[k];
        }
    }
    __syncthreads();      // the next tile reuses tile_s, cnt and tile_base
    }
}

struct SortLayout {
    size_t keys_a, vals_a, keys_b, vals_b, control, total;
    size_t control_bytes;
    int num_tiles;
};

// control block: [passes x 256 digit offsets][passes tickets (padded)][passes x tiles x 256 look-back words]
// keys per thread of the pass kernels: 8 (default: twice the tiles, 6 CTAs per SM -- the kernels are latency-bound, the
// data sits in L2) or 16 (CEBAG_SORT_ITEMS=16)
int sort_items() {
    const int items = env_int("CEBAG_SORT_ITEMS", 8) >= 16 ? 16 : 8;      // read per call
    return items;
}

SortLayout sort_layout(int64_t n) {
    SortLayout L;
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    L.num_tiles = (int)ceil_div(n > 0 ? n : 1, (int64_t)kSortThreads * 8);     // sized for the smaller tile
    size_t arr = align((size_t)(n > 0 ? n : 1) * 4);
    size_t off = 0;
    L.keys_a = off; off += arr;
    L.vals_a = off; off += arr;
    L.keys_b = off; off += arr;
    L.vals_b = off; off += arr;
    L.control_bytes = align((size_t)kMaxPasses * kDigits * 4) + align(64 * 4) +
                      align((size_t)kMaxPasses * L.num_tiles * kDigits * 4);
    L.control = off; off += L.control_bytes;
    L.total = off;
    return L;
}

}  // namespace

size_t radix_sort_workspace_bytes(int64_t n) { return sort_layout(n).total; }

void radix_sort_result(int64_t n, int key_bits, void* workspace, const uint32_t** keys_sorted,
                       const uint32_t** vals_sorted) {
    SortLayout L = sort_layout(n);
    char* ws = reinterpret_cast<char*>(workspace);
    const int passes = (key_bits + kDigitBits - 1) / kDigitBits;
    const bool in_b = ((passes - 1) & 1) != 0;
    *keys_sorted = reinterpret_cast<const uint32_t*>(ws + (in_b ? L.keys_b : L.keys_a));
    *vals_sorted = reinterpret_cast<const uint32_t*>(ws + (in_b ? L.vals_b : L.vals_a));
}

namespace {

// keys_first: int64 slot ids (first pass reads them directly) or null when the pairs already sit in the b buffers
int radix_sort_impl(const int64_t* slot_ids, int64_t n, int key_bits, void* workspace, size_t workspace_bytes,
                    const uint32_t* init_vals, const uint32_t** keys_sorted, const uint32_t** vals_sorted,
                    cudaStream_t stream) {
    CEBAG_REQUIRE(n > 0 && n < ((int64_t)1 << 30), "radix sort size");
    CEBAG_REQUIRE(key_bits >= 1 && key_bits <= 32, "radix sort key bits");
    SortLayout L = sort_layout(n);
    CEBAG_REQUIRE(workspace_bytes >= L.total, "radix sort workspace too small");
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    char* ws = reinterpret_cast<char*>(workspace);
    uint32_t* kbuf[2] = {reinterpret_cast<uint32_t*>(ws + L.keys_a), reinterpret_cast<uint32_t*>(ws + L.keys_b)};
    uint32_t* vbuf[2] = {reinterpret_cast<uint32_t*>(ws + L.vals_a), reinterpret_cast<uint32_t*>(ws + L.vals_b)};
    uint32_t* hist = reinterpret_cast<uint32_t*>(ws + L.control);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(ws + L.control + align((size_t)kMaxPasses * kDigits * 4));
    uint32_t* states = reinterpret_cast<uint32_t*>(ws + L.control + align((size_t)kMaxPasses * kDigits * 4) + align(64 * 4));
    const int passes = (key_bits + kDigitBits - 1) / kDigitBits;
    KernelScope scope(kKernSort, stream, 2 + passes);
    CEBAG_CUDA_CHECK(cudaMemsetAsync(ws + L.control, 0, L.control_bytes, stream));
    const int items = sort_items();
    const int tiles = (int)ceil_div(n, (int64_t)kSortThreads * items);
    // CTAs of the pass kernels (read per call; 0 = one per tile).  The sort runs on the look-ahead side stream next to
    // the bandwidth-bound forward / backward: one CTA per SM makes the sort itself slower (152 vs 101 us at n = 1.7 M)
    // and the step faster (0.539 vs 0.570 ms at Criteo-1TB) -- it takes fewer SM slots from kernels that need them.
    const int cta_cap = env_int("CEBAG_SORT_CTAS", kNumSMs);
    const int pass_grid = cta_cap > 0 && cta_cap < tiles ? cta_cap : tiles;
    const int64_t hcap = (int64_t)kNumSMs * env_int("CEBAG_SORT_HIST_CTAS_PER_SM", 4);
    const int hgrid = (int)(ceil_div(n, kHistTile) < hcap ? ceil_div(n, kHistTile) : hcap);
    const bool first_is_i64 = slot_ids != nullptr;
    if (first_is_i64) sort_histogram_kernel<int64_t><<<hgrid, kSortThreads, 0, stream>>>(slot_ids, n, passes, key_bits, hist);
    else sort_histogram_kernel<uint32_t><<<hgrid, kSortThreads, 0, stream>>>(kbuf[1], n, passes, key_bits, hist);
    sort_offsets_kernel<<<1, kSortThreads, 0, stream>>>(hist, passes);
    CEBAG_LAUNCH_CHECK();
    const void* kin = first_is_i64 ? static_cast<const void*>(slot_ids) : static_cast<const void*>(kbuf[1]);
    const uint32_t* vin = first_is_i64 ? init_vals : vbuf[1];
    for (int p = 0; p < passes; ++p) {
        const int shift = p * kDigitBits;
        const int bits = key_bits - shift < kDigitBits ? key_bits - shift : kDigitBits;
        uint32_t* kout = kbuf[p & 1];
        uint32_t* vout = vbuf[p & 1];
        uint32_t* state = states + (size_t)p * L.num_tiles * kDigits;
#define LAUNCH_PASS(FIRST, ITEMS)                                                                                        \
        sort_onesweep_kernel<FIRST, ITEMS><<<pass_grid, kSortThreads, 0, stream>>>(kin, vin, n, shift, bits,                \
                                                                                   hist + p * kDigits, state, tickets + p,  \
                                                                                   (uint32_t)tiles, kout, vout)
        if (p == 0 && first_is_i64) { if (items == 8) LAUNCH_PASS(true, 8); else LAUNCH_PASS(true, 16); }
        else { if (items == 8) LAUNCH_PASS(false, 8); else LAUNCH_PASS(false, 16); }
#undef LAUNCH_PASS
        CEBAG_LAUNCH_CHECK();
        kin = kout;
        vin = vout;
    }
    *keys_sorted = reinterpret_cast<const uint32_t*>(kin);
    *vals_sorted = vin;
    return CEBAG_OK;
}

}  // namespace

int radix_sort_slots(const int64_t* slot_ids, int64_t n, int key_bits, void* workspace, size_t workspace_bytes,
                     const uint32_t* init_vals, const uint32_t** keys_sorted, const uint32_t** vals_sorted,
                     cudaStream_t stream) {
    CEBAG_REQUIRE(slot_ids != nullptr, "slot ids");
    return radix_sort_impl(slot_ids, n, key_bits, workspace, workspace_bytes, init_vals, keys_sorted, vals_sorted, stream);
}

void radix_sort_input_buffers(int64_t n, void* workspace, uint32_t** keys_in, uint32_t** vals_in) {
    SortLayout L = sort_layout(n);
    char* ws = reinterpret_cast<char*>(workspace);
    *keys_in = reinterpret_cast<uint32_t*>(ws + L.keys_b);      // pass 0 reads b and writes a
    *vals_in = reinterpret_cast<uint32_t*>(ws + L.vals_b);
}

int radix_sort_u32(int64_t n, int key_bits, void* workspace, size_t workspace_bytes, const uint32_t** keys_sorted,
                   const uint32_t** vals_sorted, cudaStream_t stream) {
    return radix_sort_impl(nullptr, n, key_bits, workspace, workspace_bytes, nullptr, keys_sorted, vals_sorted, stream);
}

}  // namespace cebag
