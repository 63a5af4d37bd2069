// Probe (not product code): which engine moves a window's victims (HBM -> scattered rows of the pinned host table) and
// its missed rows (scattered rows of the pinned host table -> HBM) fastest, alone and at the same time?
//   A  zero-copy scatter kernel (STG.128), several grids
//   B  scatter through shared memory with TMA bulk stores (cp.async.bulk.global.shared::cta), several grids
//   C  gather kernel || scatter kernel variants
//   D  cudaMemcpyBatchAsync: 160 k separate 512 B copies on the copy engines, D2H, H2D, both
//   E  contiguous cudaMemcpyAsync D2H into a pinned ring + CPU threads scattering into the table (1..16 threads)
//   F  gather kernel || contiguous D2H memcpy, gather grids 56 / 148
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pcie_swap pcie_swap_probe.cu -lpthread
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>
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
// each warp: rows -> its shared-memory ring (DEPTH rows) -> one 512 B TMA bulk store per row
template <int DEPTH>
__global__ void scatter_bulk(float4* __restrict__ host, const int* __restrict__ rows, const float4* __restrict__ src, int m) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float4* ring = reinterpret_cast<float4*>(smem) + (size_t)wib * DEPTH * 32;
    long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
    int k = 0;
    for (long j = warp; j < m; j += nw, ++k) {
        const int slot = k % DEPTH;
        if (k >= DEPTH) {   // the store that used this ring entry DEPTH rows ago has read its source
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(DEPTH - 1) : "memory");
            __syncwarp();
        }
        ring[slot * 32 + lane] = src[j * 32 + lane];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            unsigned s = (unsigned)__cvta_generic_to_shared(ring + slot * 32);
            float4* d = host + (long)rows[j] * 32;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" :: "l"(d), "r"(s) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main() {
    const long N = 32L << 20;            // 32 Mi rows x 512 B = 16 GiB pinned "table"
    const int M = 160000;                // rows per direction per window (Criteo-1TB, 1 % cache)
    float4* host; CK(cudaHostAlloc(&host, N * 512, cudaHostAllocMapped | cudaHostAllocPortable));
    float4* hdev; CK(cudaHostGetDevicePointer(&hdev, host, 0));
    float4* ring; CK(cudaHostAlloc(&ring, (long)M * 512, cudaHostAllocPortable));
    memset(ring, 1, (long)M * 512);
    std::vector<int> r1(M), r2(M);
    srand(1);
    for (int i = 0; i < M; ++i) { r1[i] = (int)(((long)rand() * 65536 + rand()) % N); r2[i] = (int)(((long)rand() * 65536 + rand()) % N); }
    std::sort(r1.begin(), r1.end());
    int *d1, *d2; CK(cudaMalloc(&d1, M * 4)); CK(cudaMalloc(&d2, M * 4));
    CK(cudaMemcpy(d1, r1.data(), M * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d2, r2.data(), M * 4, cudaMemcpyHostToDevice));
    float4 *a, *b; CK(cudaMalloc(&a, (long)M * 512)); CK(cudaMalloc(&b, (long)M * 512));
    CK(cudaMemset(b, 0, (long)M * 512));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1, f0, f1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f0); cudaEventCreate(&f1);
    auto gbs = [&](float ms) { return M * 512.0 / ms / 1e6; };
    CK(cudaFuncSetAttribute(scatter_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4 * 512));
    CK(cudaFuncSetAttribute(scatter_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8 * 512));

    auto time1 = [&](const char* name, auto&& fn) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, s1); fn(s1); cudaEventRecord(e1, s1); CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("%-58s %7.3f ms %6.1f GB/s\n", name, ms, gbs(ms));
    };
    auto time2 = [&](const char* name, auto&& fa, auto&& fb) {
        float ma = 0, mb = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, s1); cudaEventRecord(f0, s2);
            fa(s1); fb(s2);
            cudaEventRecord(e1, s1); cudaEventRecord(f1, s2); CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ma, e0, e1); cudaEventElapsedTime(&mb, f0, f1);
        }
        printf("%-58s in %7.3f ms %6.1f GB/s | out %7.3f ms %6.1f GB/s\n", name, ma, gbs(ma), mb, gbs(mb));
    };
    char name[128];
    // ---- A / B: scatter alone
    for (int G : {56, 148, 296, 592}) for (int T : {128, 256}) {
        snprintf(name, sizeof name, "A scatter STG.128 %dx%d", G, T);
        time1(name, [&](cudaStream_t s) { scatter<<<G, T, 0, s>>>(hdev, d2, b, M); });
    }
    for (int G : {56, 148, 296}) {
        snprintf(name, sizeof name, "B scatter TMA bulk store depth4 %dx256", G);
        time1(name, [&](cudaStream_t s) { scatter_bulk<4><<<G, 256, 8 * 4 * 512, s>>>(hdev, d2, b, M); });
        snprintf(name, sizeof name, "B scatter TMA bulk store depth8 %dx256", G);
        time1(name, [&](cudaStream_t s) { scatter_bulk<8><<<G, 256, 8 * 8 * 512, s>>>(hdev, d2, b, M); });
    }
    for (int G : {56, 148, 296}) {
        snprintf(name, sizeof name, "  gather LDG.128 %dx128", G);
        time1(name, [&](cudaStream_t s) { gather<<<G, 128, 0, s>>>(hdev, d1, a, M); });
    }
    time1("  memcpy D2H contiguous", [&](cudaStream_t s) { CK(cudaMemcpyAsync(ring, b, (long)M * 512, cudaMemcpyDeviceToHost, s)); });
    time1("  memcpy H2D contiguous", [&](cudaStream_t s) { CK(cudaMemcpyAsync(a, ring, (long)M * 512, cudaMemcpyHostToDevice, s)); });
    // ---- C: both directions from kernels
    time2("C gather 56x128 || scatter STG 56x128",
          [&](cudaStream_t s) { gather<<<56, 128, 0, s>>>(hdev, d1, a, M); },
          [&](cudaStream_t s) { scatter<<<56, 128, 0, s>>>(hdev, d2, b, M); });
    time2("C gather 56x128 || scatter TMA depth4 56x256",
          [&](cudaStream_t s) { gather<<<56, 128, 0, s>>>(hdev, d1, a, M); },
          [&](cudaStream_t s) { scatter_bulk<4><<<56, 256, 8 * 4 * 512, s>>>(hdev, d2, b, M); });
    time2("C gather 148x128 || scatter TMA depth8 148x256",
          [&](cudaStream_t s) { gather<<<148, 128, 0, s>>>(hdev, d1, a, M); },
          [&](cudaStream_t s) { scatter_bulk<8><<<148, 256, 8 * 8 * 512, s>>>(hdev, d2, b, M); });
    // ---- F: gather kernel || contiguous D2H memcpy
    for (int G : {28, 56, 148}) {
        snprintf(name, sizeof name, "F gather %dx128 || memcpy D2H contiguous", G);
        time2(name, [&](cudaStream_t s) { gather<<<G, 128, 0, s>>>(hdev, d1, a, M); },
              [&](cudaStream_t s) { CK(cudaMemcpyAsync(ring, b, (long)M * 512, cudaMemcpyDeviceToHost, s)); });
    }
    time2("F memcpy H2D || memcpy D2H (contiguous)",
          [&](cudaStream_t s) { CK(cudaMemcpyAsync(a, ring, (long)M * 512, cudaMemcpyHostToDevice, s)); },
          [&](cudaStream_t s) { CK(cudaMemcpyAsync(ring, b, (long)M * 512, cudaMemcpyDeviceToHost, s)); });
    // ---- D: batched small copies on the copy engines
    {
        std::vector<void*> dsts(M), srcs(M);
        std::vector<size_t> sizes(M, 512);
        cudaMemcpyAttributes attr;
        memset(&attr, 0, sizeof attr);
        attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        attr.flags = cudaMemcpyFlagPreferOverlapWithCompute;
        size_t attr_idx = 0, fail = 0;
        auto batch = [&](bool d2h, cudaStream_t s, int count) {
            const std::vector<int>& rr = d2h ? r2 : r1;
            for (int i = 0; i < count; ++i) {
                if (d2h) { dsts[i] = (char*)host + (long)rr[i] * 512; srcs[i] = (char*)b + (long)i * 512; }
                else { srcs[i] = (char*)host + (long)rr[i] * 512; dsts[i] = (char*)a + (long)i * 512; }
            }
            cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), count, &attr, &attr_idx, 1, &fail, s);
            if (e != cudaSuccess) { printf("cudaMemcpyBatchAsync: %s (fail idx %zu)\n", cudaGetErrorString(e), fail); cudaGetLastError(); }
        };
        for (int count : {10000, M}) {
            for (int d2h = 0; d2h < 2; ++d2h) {
                float ms = 0; double host_ms = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0, s1);
                    double t0 = now_ms(); batch(d2h, s1, count); host_ms = now_ms() - t0;
                    cudaEventRecord(e1, s1); CK(cudaDeviceSynchronize());
                    cudaEventElapsedTime(&ms, e0, e1);
                }
                printf("D cudaMemcpyBatchAsync %s %6d x 512 B: device %7.3f ms %6.1f GB/s, host call %7.3f ms\n",
                       d2h ? "D2H" : "H2D", count, ms, count * 512.0 / ms / 1e6, host_ms);
            }
        }
    }
    // ---- E: CPU threads scatter / gather between a pinned ring and the pinned table
    for (int T : {1, 2, 4, 8, 16}) {
        for (int dir = 0; dir < 2; ++dir) {
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                double t0 = now_ms();
                std::vector<std::thread> th;
                for (int t = 0; t < T; ++t) th.emplace_back([&, t]() {
                    const long lo = (long)M * t / T, hi = (long)M * (t + 1) / T;
                    const std::vector<int>& rr = dir ? r2 : r1;
                    for (long j = lo; j < hi; ++j) {
                        if (dir) memcpy((char*)host + (long)rr[j] * 512, (char*)ring + j * 512, 512);
                        else memcpy((char*)ring + j * 512, (char*)host + (long)rr[j] * 512, 512);
                    }
                });
                for (auto& x : th) x.join();
                best = std::min(best, now_ms() - t0);
            }
            printf("E CPU %s %2d threads (incl. spawn): %7.3f ms %6.1f GB/s\n", dir ? "scatter ring->table" : "gather table->ring",
                   T, best, M * 512.0 / best / 1e6);
        }
    }
    printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
    return 0;
}
