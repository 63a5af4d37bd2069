// Pinned host memory for the table, error reporting, and the counter-based uniform fill.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "profile.cuh"

namespace cebag {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

namespace {

// splitmix64 finaliser: element i of stream `seed` depends only on (seed, i)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ float uniform_at(uint64_t seed, int64_t i, float lo, float span) {
    uint32_t bits = (uint32_t)(mix64(seed ^ mix64((uint64_t)i)) >> 40);   // 24 random bits
    return lo + span * ((float)bits * (1.0f / 16777216.0f));
}

__global__ void __launch_bounds__(256)
fill_uniform_kernel(float* __restrict__ dst, int64_t count, float lo, float span, uint64_t seed, int vec) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (vec) {
        int64_t n4 = count / 4;
        for (int64_t q = tid; q < n4; q += stride) {
            float4 v = make_float4(uniform_at(seed, 4 * q, lo, span), uniform_at(seed, 4 * q + 1, lo, span),
                                   uniform_at(seed, 4 * q + 2, lo, span), uniform_at(seed, 4 * q + 3, lo, span));
            reinterpret_cast<float4*>(dst)[q] = v;
        }
        for (int64_t i = n4 * 4 + tid; i < count; i += stride) dst[i] = uniform_at(seed, i, lo, span);
    } else {
        for (int64_t i = tid; i < count; i += stride) dst[i] = uniform_at(seed, i, lo, span);
    }
}

}  // namespace
}  // namespace cebag

using namespace cebag;

extern "C" int cebag_abi_version(void) { return CEBAG_ABI_VERSION; }

extern "C" const char* cebag_last_error(void) { return g_error; }

extern "C" int cebag_host_alloc(void** out_ptr, size_t bytes) {
    CEBAG_REQUIRE(out_ptr != nullptr && bytes > 0, "host_alloc arguments");
    CEBAG_CUDA_CHECK(cudaHostAlloc(out_ptr, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    return CEBAG_OK;
}

extern "C" int cebag_host_free(void* ptr) {
    if (ptr) CEBAG_CUDA_CHECK(cudaFreeHost(ptr));
    return CEBAG_OK;
}

extern "C" int cebag_host_register(void* ptr, size_t bytes) {
    CEBAG_REQUIRE(ptr != nullptr && bytes > 0, "host_register arguments");
    CEBAG_CUDA_CHECK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return CEBAG_OK;
}

extern "C" int cebag_host_unregister(void* ptr) {
    if (ptr) CEBAG_CUDA_CHECK(cudaHostUnregister(ptr));
    return CEBAG_OK;
}

extern "C" int cebag_host_device_pointer(void* host_ptr, void** out_dev_ptr) {
    CEBAG_REQUIRE(host_ptr != nullptr && out_dev_ptr != nullptr, "host_device_pointer arguments");
    CEBAG_CUDA_CHECK(cudaHostGetDevicePointer(out_dev_ptr, host_ptr, 0));
    return CEBAG_OK;
}

extern "C" int cebag_device_alloc(void** out_ptr, size_t bytes) {
    CEBAG_REQUIRE(out_ptr != nullptr && bytes > 0, "device_alloc arguments");
    CEBAG_CUDA_CHECK(cudaMalloc(out_ptr, bytes));
    return CEBAG_OK;
}

extern "C" int cebag_device_free(void* ptr) {
    if (ptr) CEBAG_CUDA_CHECK(cudaFree(ptr));
    return CEBAG_OK;
}

extern "C" int cebag_ipc_export(void* ptr, unsigned char handle[64]) {
    CEBAG_REQUIRE(ptr != nullptr && handle != nullptr, "ipc_export arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CEBAG_CUDA_CHECK(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, 64);
    return CEBAG_OK;
}

extern "C" int cebag_ipc_import(const unsigned char handle[64], void** out_ptr) {
    CEBAG_REQUIRE(handle != nullptr && out_ptr != nullptr, "ipc_import arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CEBAG_CUDA_CHECK(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CEBAG_OK;
}

extern "C" int cebag_ipc_close(void* ptr) {
    if (ptr) CEBAG_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return CEBAG_OK;
}

extern "C" int cebag_fill_uniform(float* dst, int64_t count, float lo, float hi, uint64_t seed, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(dst != nullptr && count >= 0, "fill_uniform arguments");
    if (count == 0) return CEBAG_OK;
    KernelScope scope(kKernFill, stream);
    fill_uniform_kernel<<<kNumSMs * 8, 256, 0, stream>>>(dst, count, lo, hi - lo, seed, aligned16(dst) ? 1 : 0);
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}

// ---- stream-ordered barrier over peer memory -----------------------------------------------------------------------------
// Every rank owns an array of CEBAG_MAX_PEERS 32-bit flags that all peers have mapped (CUDA IPC).  Round `seq` (a device
// counter that every barrier advances): rank r stores seq into flags_of_peer[j][r] for every j (a release store over NVLink), then waits until its own flags[j] >= seq
// for every j.  One CTA, one thread per peer: ~3 us, no NCCL call and no host involvement, so the fused exchange's two
// barriers per step cost two tiny kernels.  A peer that never arrives trips the timeout instead of hanging the GPU.
namespace cebag {
namespace {
__global__ void peer_barrier_kernel(cebag_exchange x, int rank, uint32_t* seq_counter, long long timeout_cycles,
                                    int32_t* failed) {
    // the round number lives in device memory and advances with every barrier, so a captured CUDA graph that contains
    // this kernel can be replayed (every rank executes the same sequence of barriers)
    __shared__ uint32_t seq_s;
    if (threadIdx.x == 0) seq_s = *seq_counter + 1u;
    __syncthreads();
    const uint32_t seq = seq_s;
    if (threadIdx.x == 0) *seq_counter = seq;
    const int j = threadIdx.x;
    if (j >= x.world) return;
    volatile uint32_t* mine = reinterpret_cast<volatile uint32_t*>(x.peer[rank]);
    uint32_t* theirs = reinterpret_cast<uint32_t*>(x.peer[j]);
    __threadfence_system();                                   // my earlier stores (pooled rows) are visible first
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(theirs + rank), "r"(seq) : "memory");
    const long long t0 = clock64();
    while ((int32_t)(mine[j] - seq) < 0) {
        if (clock64() - t0 > timeout_cycles) { *failed = 1; break; }
    }
    __threadfence_system();
}
}  // namespace
}  // namespace cebag

extern "C" int cebag_peer_barrier(const cebag_exchange* flags, int32_t rank, uint32_t* seq_counter, int32_t* failed_flag,
                                  void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CEBAG_REQUIRE(flags != nullptr && failed_flag != nullptr && seq_counter != nullptr, "barrier arguments");
    CEBAG_REQUIRE(flags->world >= 1 && flags->world <= CEBAG_MAX_PEERS && rank >= 0 && rank < flags->world, "barrier ranks");
    for (int q = 0; q < flags->world; ++q) CEBAG_REQUIRE(flags->peer[q] != nullptr, "barrier flag pointer");
    count_launches(1);
    // ~4 s at 1.9 GHz: far beyond any skew between ranks of one step, short of the driver's watchdogs
    cebag::peer_barrier_kernel<<<1, 32, 0, stream>>>(*flags, rank, seq_counter, 8000000000LL, failed_flag);
    CEBAG_LAUNCH_CHECK();
    return CEBAG_OK;
}
