// Stable LSD radix sort of (slot, lookup position) pairs -- the integer preprocessing of the fused backward.
// Keys are slot ids (< C, so only ceil(log2 C) bits are sorted); values are the original lookup positions.
// Per pass: tile histogram -> exclusive scan of the bucket-major histogram -> stable scatter (match.any ranking).
#include "common.cuh"
#include "scan.cuh"
#include "radix_sort.cuh"
#include "profile.cuh"

namespace cebag {

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;                               // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;         // 4096 keys per CTA
constexpr int kMaxDigitBits = 11;                           // 21-bit slot ids sort in 2 passes
constexpr int kMaxBuckets = 1 << kMaxDigitBits;

template <bool FIRST>
__device__ __forceinline__ uint32_t load_key(const void* keys_in, int64_t i) {
    if (FIRST) return (uint32_t) reinterpret_cast<const int64_t*>(keys_in)[i];
    return reinterpret_cast<const uint32_t*>(keys_in)[i];
}

// tile histogram, written bucket-major: hist[b * num_tiles + tile]
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const void* __restrict__ keys_in, int64_t n, int shift, int nbuckets, int num_tiles,
                 int32_t* __restrict__ hist) {
    __shared__ int32_t h[kMaxBuckets];
    for (int b = threadIdx.x; b < nbuckets; b += kSortThreads) h[b] = 0;
    __syncthreads();
    const uint32_t mask = (uint32_t)nbuckets - 1u;
    int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        int64_t i = base + (int64_t)k * kSortThreads + threadIdx.x;
        bool valid = i < n;
        uint32_t d = valid ? ((load_key<FIRST>(keys_in, i) >> shift) & mask) : (0x80000000u | (uint32_t)lane_id());
        // a warp whose 32 keys share the digit (hot slots of small tables) adds once; otherwise plain shared atomics
        int same = 0;
        __match_all_sync(0xffffffffu, d, &same);
        if (same) { if (lane_id() == 0 && valid) atomicAdd(&h[d], 32); }
        else if (valid) atomicAdd(&h[d], 1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbuckets; b += kSortThreads) hist[(int64_t)b * num_tiles + blockIdx.x] = h[b];
}

// stable scatter of one tile.  Tile order = warp-major, then round-major, then lane: position within the tile is
// warp * (32 * ITEMS) + round * 32 + lane, which the ranking below preserves inside every bucket.
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads)
sort_scatter_kernel(const void* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, int64_t n, int shift,
                    int nbuckets, int num_tiles, const int32_t* __restrict__ hist_scanned,
                    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint16_t cnt[kSortWarps][kMaxBuckets];   // per-warp bucket counts (<= 32 * kSortItems = 512)
    __shared__ int32_t gbase[kMaxBuckets];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    for (int b = threadIdx.x; b < kSortWarps * nbuckets; b += kSortThreads) cnt[b / nbuckets][b % nbuckets] = 0;
    __syncthreads();
    const uint32_t mask = (uint32_t)nbuckets - 1u;
    const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (32 * kSortItems);
    uint32_t key[kSortItems], val[kSortItems];
    int32_t rank[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        int64_t i = wbase + k * 32 + lane;
        bool valid = i < n;
        key[k] = valid ? load_key<FIRST>(keys_in, i) : 0u;
        val[k] = valid ? (vals_in ? vals_in[i] : (uint32_t)i) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        int64_t i = wbase + k * 32 + lane;
        bool valid = i < n;
        uint32_t d = valid ? ((key[k] >> shift) & mask) : (0x80000000u | (uint32_t)lane);  // invalid: match only self
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        int32_t old = 0;
        if (valid && lane == leader) {
            old = cnt[warp][d];
            cnt[warp][d] = (uint16_t)(old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = old + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbuckets; b += kSortThreads) {
        int32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            int32_t c = cnt[w][b];
            cnt[w][b] = (uint16_t)run;
            run += c;
        }
        gbase[b] = hist_scanned[(int64_t)b * num_tiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        int64_t i = wbase + k * 32 + lane;
        if (i < n) {
            uint32_t d = (key[k] >> shift) & mask;
            int64_t pos = (int64_t)gbase[d] + cnt[warp][d] + rank[k];
            keys_out[pos] = key[k];
            vals_out[pos] = val[k];
        }
    }
}

struct SortLayout {
    size_t keys_a, vals_a, keys_b, vals_b, hist, scan_ws, total;
    int num_tiles;
};

SortLayout sort_layout(int64_t n) {
    SortLayout L;
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    L.num_tiles = (int)ceil_div(n > 0 ? n : 1, kSortTile);
    size_t arr = align((size_t)(n > 0 ? n : 1) * 4);
    size_t off = 0;
    L.keys_a = off; off += arr;
    L.vals_a = off; off += arr;
    L.keys_b = off; off += arr;
    L.vals_b = off; off += arr;
    L.hist = off; off += align((size_t)kMaxBuckets * L.num_tiles * 4);
    L.scan_ws = off; off += align(scan_workspace_bytes((int64_t)kMaxBuckets * L.num_tiles));
    L.total = off;
    return L;
}

}  // namespace

size_t radix_sort_workspace_bytes(int64_t n) { return sort_layout(n).total; }

void radix_sort_result(int64_t n, int key_bits, void* workspace, const uint32_t** keys_sorted,
                       const uint32_t** vals_sorted) {
    SortLayout L = sort_layout(n);
    char* ws = reinterpret_cast<char*>(workspace);
    const int passes = (key_bits + kMaxDigitBits - 1) / kMaxDigitBits;
    const bool in_b = ((passes - 1) & 1) != 0;
    *keys_sorted = reinterpret_cast<const uint32_t*>(ws + (in_b ? L.keys_b : L.keys_a));
    *vals_sorted = reinterpret_cast<const uint32_t*>(ws + (in_b ? L.vals_b : L.vals_a));
}

int radix_sort_slots(const int64_t* slot_ids, int64_t n, int key_bits, void* workspace, size_t workspace_bytes,
                     const uint32_t* init_vals, const uint32_t** keys_sorted, const uint32_t** vals_sorted,
                     cudaStream_t stream) {
    CEBAG_REQUIRE(n > 0 && n < ((int64_t)1 << 31), "radix sort size");
    CEBAG_REQUIRE(key_bits >= 1 && key_bits <= 32, "radix sort key bits");
    SortLayout L = sort_layout(n);
    CEBAG_REQUIRE(workspace_bytes >= L.total, "radix sort workspace too small");
    char* ws = reinterpret_cast<char*>(workspace);
    uint32_t* kbuf[2] = {reinterpret_cast<uint32_t*>(ws + L.keys_a), reinterpret_cast<uint32_t*>(ws + L.keys_b)};
    uint32_t* vbuf[2] = {reinterpret_cast<uint32_t*>(ws + L.vals_a), reinterpret_cast<uint32_t*>(ws + L.vals_b)};
    int32_t* hist = reinterpret_cast<int32_t*>(ws + L.hist);
    int32_t* scan_ws = reinterpret_cast<int32_t*>(ws + L.scan_ws);
    const int passes = (key_bits + kMaxDigitBits - 1) / kMaxDigitBits;
    KernelScope scope(kKernSort, stream, 2 * passes);
    const int bits_per_pass = (key_bits + passes - 1) / passes;
    int shift = 0;
    const void* kin = slot_ids;
    const uint32_t* vin = init_vals;
    for (int p = 0; p < passes; ++p) {
        int bits = (key_bits - shift) < bits_per_pass ? (key_bits - shift) : bits_per_pass;
        int nbuckets = 1 << bits;
        uint32_t* kout = kbuf[p & 1];
        uint32_t* vout = vbuf[p & 1];
        if (p == 0) {
            sort_hist_kernel<true><<<L.num_tiles, kSortThreads, 0, stream>>>(kin, n, shift, nbuckets, L.num_tiles, hist);
        } else {
            sort_hist_kernel<false><<<L.num_tiles, kSortThreads, 0, stream>>>(kin, n, shift, nbuckets, L.num_tiles, hist);
        }
        CEBAG_LAUNCH_CHECK();
        int rc = exclusive_scan_inplace(hist, (int64_t)nbuckets * L.num_tiles, nullptr, scan_ws, stream);
        if (rc) return rc;
        if (p == 0) {
            sort_scatter_kernel<true><<<L.num_tiles, kSortThreads, 0, stream>>>(kin, vin, n, shift, nbuckets,
                                                                                L.num_tiles, hist, kout, vout);
        } else {
            sort_scatter_kernel<false><<<L.num_tiles, kSortThreads, 0, stream>>>(kin, vin, n, shift, nbuckets,
                                                                                 L.num_tiles, hist, kout, vout);
        }
        CEBAG_LAUNCH_CHECK();
        kin = kout;
        vin = vout;
        shift += bits;
    }
    *keys_sorted = reinterpret_cast<const uint32_t*>(kin);
    *vals_sorted = vin;
    return CEBAG_OK;
}

}  // namespace cebag
