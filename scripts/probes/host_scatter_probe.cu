// Probe (not product code): how fast can host threads scatter 160 k rows of 512 B from a pinned ring into random rows of
// the pinned table, (a) with cudaHostAlloc memory vs transparent-huge-page backed + cudaHostRegister'ed memory,
// (b) with memcpy / software prefetch / non-temporal AVX-512 stores, (c) with sorted destinations.
// Also re-times the zero-copy gather / scatter kernels on the huge-page table.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -mavx512f -o host_scatter host_scatter_probe.cu -lpthread
#include <cuda_runtime.h>
#include <immintrin.h>
#include <sys/mman.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>
#include <atomic>
#include <chrono>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void gather(const float4* __restrict__ host, const int* __restrict__ rows, float4* __restrict__ dst, int m) {
    int lane = threadIdx.x & 31;
    long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
    for (long j = warp; j < m; j += nw) dst[j * 32 + lane] = host[(long)rows[j] * 32 + lane];
}
__global__ void scatter(float4* __restrict__ host, const int* __restrict__ rows, const float4* __restrict__ src, int m) {
    int lane = threadIdx.x & 31;
    long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
    for (long j = warp; j < m; j += nw) host[(long)rows[j] * 32 + lane] = src[j * 32 + lane];
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static inline void row_copy_nt(char* dst, const char* src) {
#pragma unroll
    for (int k = 0; k < 8; ++k) _mm512_stream_si512((__m512i*)(dst + 64 * k), _mm512_load_si512((const __m512i*)(src + 64 * k)));
}

enum { kMemcpy = 0, kPrefetch = 1, kNT = 2, kNTPrefetchSrc = 3 };
static const char* kNames[] = {"memcpy", "prefetchw+memcpy", "AVX-512 NT stores", "NT stores + TLB prefetch"};

static void scatter_cpu(char* table, const char* ring, const int* rows, long lo, long hi, int variant) {
    const int kAhead = 12;
    for (long j = lo; j < hi; ++j) {
        char* dst = table + (long)rows[j] * 512;
        if (variant == kPrefetch && j + kAhead < hi) {
            char* p = table + (long)rows[j + kAhead] * 512;
            for (int k = 0; k < 8; ++k) __builtin_prefetch(p + 64 * k, 1, 3);
        }
        if (variant == kNTPrefetchSrc && j + kAhead < hi) __builtin_prefetch(table + (long)rows[j + kAhead] * 512, 0, 0);
        if (variant == kMemcpy || variant == kPrefetch) memcpy(dst, ring + j * 512, 512);
        else row_copy_nt(dst, ring + j * 512);
    }
    _mm_sfence();
}

int main() {
    const long N = 32L << 20;
    const long BYTES = N * 512;
    const int M = 160000, REPS = 4;
    FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    if (f) { char buf[128] = {0}; fgets(buf, 127, f); printf("THP enabled: %s", buf); fclose(f); }
    char* tabA; CK(cudaHostAlloc(&tabA, BYTES, cudaHostAllocMapped | cudaHostAllocPortable));
    double t0 = now_ms();
    char* tabB = (char*)mmap(nullptr, BYTES + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (tabB == MAP_FAILED) { printf("mmap failed\n"); return 1; }
    tabB = (char*)(((uintptr_t)tabB + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
    int adv = madvise(tabB, BYTES, MADV_HUGEPAGE);
    {   // first touch in parallel
        std::vector<std::thread> th;
        for (int t = 0; t < 16; ++t) th.emplace_back([&, t]() { memset(tabB + BYTES / 16 * t, 0, BYTES / 16); });
        for (auto& x : th) x.join();
    }
    double t1 = now_ms();
    CK(cudaHostRegister(tabB, BYTES, cudaHostRegisterPortable | cudaHostRegisterMapped));
    double t2 = now_ms();
    printf("huge-page table: madvise rc %d, touch %.0f ms, cudaHostRegister %.0f ms\n", adv, t1 - t0, t2 - t1);
    f = fopen("/proc/meminfo", "r");
    if (f) { char line[256]; while (fgets(line, 255, f)) if (strstr(line, "AnonHugePages") || strstr(line, "Hugepagesize")) printf("%s", line); fclose(f); }

    char* ring; CK(cudaHostAlloc(&ring, (long)M * 512 * REPS, cudaHostAllocPortable));
    memset(ring, 1, (long)M * 512 * REPS);
    std::vector<int> rows((size_t)M * REPS), rows_sorted;
    srand(1);
    for (auto& r : rows) r = (int)(((long)rand() * 65536 + rand()) % N);
    rows_sorted = rows;
    for (int r = 0; r < REPS; ++r) std::sort(rows_sorted.begin() + (size_t)r * M, rows_sorted.begin() + (size_t)(r + 1) * M);

    for (int which = 0; which < 2; ++which) {
        char* table = which ? tabB : tabA;
        printf("---- table in %s\n", which ? "THP + cudaHostRegister" : "cudaHostAlloc");
        for (int sorted = 0; sorted < 2; ++sorted) {
            const int* rr = sorted ? rows_sorted.data() : rows.data();
            for (int variant = 0; variant < 4; ++variant) {
                for (int T : {1, 2, 4, 8}) {
                    std::atomic<int> ready{0};
                    std::atomic<bool> go{false};
                    double tstart = 0;
                    std::vector<std::thread> th;
                    for (int t = 0; t < T; ++t) th.emplace_back([&, t]() {
                        ready.fetch_add(1);
                        while (!go.load(std::memory_order_acquire)) {}
                        for (int r = 0; r < REPS; ++r) {
                            const long lo = (long)r * M + (long)M * t / T, hi = (long)r * M + (long)M * (t + 1) / T;
                            scatter_cpu(table, ring, rr, lo, hi, variant);
                        }
                    });
                    while (ready.load() < T) {}
                    tstart = now_ms();
                    go.store(true, std::memory_order_release);
                    for (auto& x : th) x.join();
                    double ms = (now_ms() - tstart) / REPS;
                    printf("CPU scatter %-26s %s %d threads: %7.3f ms per 160k rows, %5.1f GB/s\n", kNames[variant],
                           sorted ? "sorted  " : "unsorted", T, ms, M * 512.0 / ms / 1e6);
                }
            }
        }
    }
    // zero-copy kernels on both tables
    int* d_rows; CK(cudaMalloc(&d_rows, M * 4)); CK(cudaMemcpy(d_rows, rows_sorted.data(), M * 4, cudaMemcpyHostToDevice));
    float4* dbuf; CK(cudaMalloc(&dbuf, (long)M * 512)); CK(cudaMemset(dbuf, 0, (long)M * 512));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        float4* hdev; CK(cudaHostGetDevicePointer(&hdev, which ? tabB : tabA, 0));
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); gather<<<56, 128>>>(hdev, d_rows, dbuf, M); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); }
        printf("GPU gather 56x128 on %s: %.3f ms %.1f GB/s\n", which ? "THP table" : "cudaHostAlloc table", ms, M * 512.0 / ms / 1e6);
        for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); scatter<<<56, 128>>>(hdev, d_rows, dbuf, M); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); }
        printf("GPU scatter 56x128 on %s: %.3f ms %.1f GB/s\n", which ? "THP table" : "cudaHostAlloc table", ms, M * 512.0 / ms / 1e6);
    }
    return 0;
}
