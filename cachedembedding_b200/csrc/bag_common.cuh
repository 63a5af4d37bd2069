// Row-vector helpers shared by the embedding-bag forward / backward kernels.
//
// A "group" of LANES threads (4, 8, 16 or 32; a power of two, so groups never straddle a warp) owns one row at a
// time.  Rows are D floats; with VT = float4 a lane touches 128-bit chunks (the fast path: D % 4 == 0 and 16-byte
// aligned bases), with VT = float single floats (any D).  CPL = chunks per lane, so a row is LANES * CPL chunks
// wide at most and lives entirely in registers.
#pragma once
#include "common.cuh"

namespace cebag {

struct BagParams {
    const float*   cache;
    const int64_t* slot_ids;
    const void*    offsets;
    const float*   psw;
    int64_t n;
    int64_t num_bags;
    int64_t padding_idx;
    int64_t layout_batch;     // B (sample-major layout)
    int64_t layout_features;  // F = num_bags / B
    int32_t dim;              // D floats
    int32_t cache_rows;       // C
    uint32_t key_mask;        // backward: slot = sorted key & key_mask (window plans carry the batch index above it)
    int32_t chunks;           // row width in VT chunks (D/4 or D)
    int32_t offsets_are_64;
    int32_t include_last;
    int32_t mode;
    int32_t layout;
    // CEBAG_LAYOUT_EXCHANGE: rows of out / grad_out live in the peer buffers (fp32[B_j, F, D] on rank j)
    int32_t exch_world;
    int32_t exch_feature_offset;
    int32_t exch_total_features;
    int32_t exch_base;            // B / world
    int32_t exch_rem;             // B % world: the first exch_rem ranks own exch_base + 1 samples
    float*  exch_peer[CEBAG_MAX_PEERS];
};

__device__ __forceinline__ int64_t load_offset(const BagParams& p, int64_t g) {
    // offsets[g]; the implied last offset (include_last_offset == False) is n
    if (g >= p.num_bags && !p.include_last) return p.n;
    return p.offsets_are_64 ? reinterpret_cast<const int64_t*>(p.offsets)[g]
                            : (int64_t) reinterpret_cast<const int32_t*>(p.offsets)[g];
}

// row of out / grad_out that belongs to bag g
__device__ __forceinline__ int64_t bag_row(const BagParams& p, int64_t g) {
    if (p.layout == CEBAG_LAYOUT_SAMPLE_MAJOR) {
        int64_t f = g / p.layout_batch, b = g - f * p.layout_batch;
        return b * p.layout_features + f;
    }
    return g;
}

// address of the out / grad_out row of bag g; `local` is the caller's own out / grad_out buffer
__device__ __forceinline__ float* bag_row_ptr(const BagParams& p, float* local, int64_t g) {
    if (p.layout == CEBAG_LAYOUT_EXCHANGE) {
        const int64_t f = g / p.layout_batch, b = g - f * p.layout_batch;
        const int64_t big = (int64_t)p.exch_rem * (p.exch_base + 1);      // samples held by the ranks with one extra
        int j;
        int64_t b_local;
        if (b < big) { j = (int)(b / (p.exch_base + 1)); b_local = b - (int64_t)j * (p.exch_base + 1); }
        else { j = p.exch_rem + (int)((b - big) / p.exch_base); b_local = b - big - (int64_t)(j - p.exch_rem) * p.exch_base; }
        float* base = p.exch_peer[0];
#pragma unroll
        for (int q = 1; q < CEBAG_MAX_PEERS; ++q) base = (j == q) ? p.exch_peer[q] : base;
        return base + (b_local * p.exch_total_features + p.exch_feature_offset + f) * p.dim;
    }
    return local + bag_row(p, g) * p.dim;
}

__device__ __forceinline__ const float* shfl_ptr(unsigned mask, const float* ptr, int src, int width) {
    unsigned long long v = reinterpret_cast<unsigned long long>(ptr);
    unsigned lo = __shfl_sync(mask, (unsigned)(v & 0xffffffffu), src, width);
    unsigned hi = __shfl_sync(mask, (unsigned)(v >> 32), src, width);
    return reinterpret_cast<const float*>(((unsigned long long)hi << 32) | lo);
}

template <typename VT> struct Vec;
template <> struct Vec<float4> {
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ float4 ld_stream(const float4* p) { return ld_stream_f4(p); }
    // cache row for the forward gather; POLICY 0: bypass L1, 1: allocate in L1, 2: allocate and evict last
    template <int POLICY> static __device__ __forceinline__ float4 ld_row(const float4* p) {
        return POLICY == 0 ? ld_stream_f4(p) : POLICY == 1 ? ld_nc_f4(p) : ld_nc_keep_f4(p);
    }
    static __device__ __forceinline__ float4 ld(const float4* p) { return ld_f4(p); }
    static __device__ __forceinline__ void st(float4* p, const float4& v) { st_f4(p, v); }
    static __device__ __forceinline__ void st_stream(float4* p, const float4& v) { st_stream_f4(p, v); }
    static __device__ __forceinline__ void fma(float4& a, float w, const float4& v) { fma4(a, w, v); }
    static __device__ __forceinline__ float dot(const float4& a, const float4& b) {
        return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    static __device__ __forceinline__ float4 scale(const float4& a, float s) {
        return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
    }
    // a - s * b, with the product rounded first (matches W.add_(g, alpha=-lr) on a coalesced grad)
    static __device__ __forceinline__ float4 sub_scaled(const float4& a, float s, const float4& b) {
        return make_float4(a.x - __fmul_rn(s, b.x), a.y - __fmul_rn(s, b.y), a.z - __fmul_rn(s, b.z),
                           a.w - __fmul_rn(s, b.w));
    }
};
template <> struct Vec<float> {
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float ld_stream(const float* p) { return __ldg(p); }
    template <int POLICY> static __device__ __forceinline__ float ld_row(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, const float& v) { *p = v; }
    static __device__ __forceinline__ void st_stream(float* p, const float& v) { *p = v; }
    static __device__ __forceinline__ void fma(float& a, float w, const float& v) { a = fmaf(w, v, a); }
    static __device__ __forceinline__ float dot(const float& a, const float& b) { return a * b; }
    static __device__ __forceinline__ float scale(const float& a, float s) { return a * s; }
    static __device__ __forceinline__ float sub_scaled(const float& a, float s, const float& b) {
        return a - __fmul_rn(s, b);
    }
};

// sum over the LANES threads of a group; only that group's lanes take part (groups of one warp may diverge)
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
    const unsigned mask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (lane_id() & ~(LANES - 1)));
#pragma unroll
    for (int d = LANES / 2; d > 0; d >>= 1) v += __shfl_xor_sync(mask, v, d, LANES);
    return v;
}

// Host-side dispatch over (VT, LANES, CPL).  `chunks` is the row width in VT units.
struct RowShape {
    bool vec;      // float4 path
    int  lanes;    // group size
    int  cpl;      // chunks per lane (1, 2 or 4); 0 = row too wide for the register-resident kernels
    int  chunks;
};

static inline RowShape row_shape(int dim, bool all_aligned16) {
    RowShape r;
    r.vec = (dim % 4 == 0) && all_aligned16;
    r.chunks = r.vec ? dim / 4 : dim;
    r.lanes = r.chunks <= 4 ? 4 : r.chunks <= 8 ? 8 : r.chunks <= 16 ? 16 : 32;
    int cpl = (r.chunks + r.lanes - 1) / r.lanes;
    r.cpl = cpl <= 1 ? 1 : cpl <= 2 ? 2 : cpl <= 4 ? 4 : 0;
    return r;
}

// Expands to a switch that calls MACRO(VT, LANES, CPL) for the shape `rs`.
#define CEBAG_DISPATCH_ROW_SHAPE(rs, MACRO)                                               \
    do {                                                                                  \
        if ((rs).vec) {                                                                   \
            switch ((rs).lanes * 8 + (rs).cpl) {                                          \
                case 4 * 8 + 1:  MACRO(float4, 4, 1); break;                              \
                case 8 * 8 + 1:  MACRO(float4, 8, 1); break;                              \
                case 16 * 8 + 1: MACRO(float4, 16, 1); break;                             \
                case 32 * 8 + 1: MACRO(float4, 32, 1); break;                             \
                case 32 * 8 + 2: MACRO(float4, 32, 2); break;                             \
                case 32 * 8 + 4: MACRO(float4, 32, 4); break;                             \
                default: cebag::set_error("unsupported row width %d", (rs).chunks); return CEBAG_ERR_INVALID; \
            }                                                                             \
        } else {                                                                          \
            switch ((rs).lanes * 8 + (rs).cpl) {                                          \
                case 4 * 8 + 1:  MACRO(float, 4, 1); break;                               \
                case 8 * 8 + 1:  MACRO(float, 8, 1); break;                               \
                case 16 * 8 + 1: MACRO(float, 16, 1); break;                              \
                case 32 * 8 + 1: MACRO(float, 32, 1); break;                              \
                case 32 * 8 + 2: MACRO(float, 32, 2); break;                              \
                case 32 * 8 + 4: MACRO(float, 32, 4); break;                              \
                default: cebag::set_error("unsupported row width %d", (rs).chunks); return CEBAG_ERR_INVALID; \
            }                                                                             \
        }                                                                                 \
    } while (0)

int fill_bag_params(const cebag_bag_args* a, BagParams* p, const RowShape& rs);

// TMA-driven forward (bag_forward_tma.cu): returns true when it took the call (*rc = its status)
bool bag_forward_tma_launch(const cebag_bag_args* a, const BagParams& p, float* out, cudaStream_t stream, int* rc);

}  // namespace cebag
