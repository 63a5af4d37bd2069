// Probe (not product code): do the zero-copy gather / scatter rates depend on the SIZE of the pinned table (IOMMU / GPU
// TLB reach) and on how it was allocated?  Table of `gib` GiB (default 85), 185 k rows of 512 B per direction, rows
// ascending.  Usage: pcie_bigtable [gib] [thp]
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pcie_bigtable pcie_bigtable_probe.cu -lpthread
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>
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

int main(int argc, char** argv) {
    const long gib = argc > 1 ? atol(argv[1]) : 85;
    const bool thp = argc > 2 && !strcmp(argv[2], "thp");
    const long N = gib * (1L << 30) / 512;
    const long BYTES = N * 512;
    const int M = 185000;
    char* tab;
    if (thp) {
        tab = (char*)mmap(nullptr, BYTES + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        tab = (char*)(((uintptr_t)tab + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
        madvise(tab, BYTES, MADV_HUGEPAGE);
        std::vector<std::thread> th;
        for (int t = 0; t < 16; ++t) th.emplace_back([&, t]() { memset(tab + BYTES / 16 * t, 0, BYTES / 16); });
        for (auto& x : th) x.join();
        CK(cudaHostRegister(tab, BYTES, cudaHostRegisterPortable | cudaHostRegisterMapped));
    } else {
        CK(cudaHostAlloc(&tab, BYTES, cudaHostAllocMapped | cudaHostAllocPortable));
    }
    float4* hdev; CK(cudaHostGetDevicePointer(&hdev, tab, 0));
    float4* ring; CK(cudaHostAlloc(&ring, (long)M * 512, cudaHostAllocPortable));
    printf("table %ld GiB (%s), %d rows per direction\n", gib, thp ? "THP + cudaHostRegister" : "cudaHostAlloc", M);
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1, f0, f1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f0); cudaEventCreate(&f1);
    float4 *a, *b; CK(cudaMalloc(&a, (long)M * 512)); CK(cudaMalloc(&b, (long)M * 512)); CK(cudaMemset(b, 0, (long)M * 512));
    int *d1, *d2; CK(cudaMalloc(&d1, M * 4)); CK(cudaMalloc(&d2, M * 4));
    auto gbs = [&](float ms) { return M * 512.0 / ms / 1e6; };
    // span: rows drawn from the first `span` fraction of the table (locality of the touched range), always ascending
    for (double span : {1.0, 0.18, 0.02}) {
        std::vector<int> r1(M), r2(M);
        srand(7);
        const long lim = (long)(N * span);
        for (int i = 0; i < M; ++i) { r1[i] = (int)(((long)rand() * 65536 + rand()) % lim); r2[i] = (int)(((long)rand() * 65536 + rand()) % lim); }
        std::sort(r1.begin(), r1.end());
        std::vector<int> r2u = r2;
        std::sort(r2.begin(), r2.end());
        CK(cudaMemcpy(d1, r1.data(), M * 4, cudaMemcpyHostToDevice));
        for (int sorted = 1; sorted >= 0; --sorted) {
            CK(cudaMemcpy(d2, (sorted ? r2 : r2u).data(), M * 4, cudaMemcpyHostToDevice));
            float ms = 0, ms2 = 0;
            for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0, s1); gather<<<56, 128, 0, s1>>>(hdev, sorted ? d1 : d2, a, M); cudaEventRecord(e1, s1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); }
            printf("span %.2f %s gather 56x128          %7.3f ms %5.1f GB/s\n", span, sorted ? "sorted  " : "unsorted", ms, gbs(ms));
            for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0, s1); scatter<<<56, 128, 0, s1>>>(hdev, d2, b, M); cudaEventRecord(e1, s1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); }
            printf("span %.2f %s scatter 56x128         %7.3f ms %5.1f GB/s\n", span, sorted ? "sorted  " : "unsorted", ms, gbs(ms));
        }
        float ms = 0, ms2 = 0;
        for (int G : {56, 148}) {
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0, s1); cudaEventRecord(f0, s2);
                gather<<<G, 128, 0, s1>>>(hdev, d1, a, M); CK(cudaMemcpyAsync(ring, b, (long)M * 512, cudaMemcpyDeviceToHost, s2));
                cudaEventRecord(e1, s1); cudaEventRecord(f1, s2); CK(cudaDeviceSynchronize());
                cudaEventElapsedTime(&ms, e0, e1); cudaEventElapsedTime(&ms2, f0, f1);
            }
            printf("span %.2f gather %dx128 || memcpy D2H: gather %7.3f ms %5.1f GB/s | memcpy %7.3f ms %5.1f GB/s\n", span, G, ms, gbs(ms), ms2, gbs(ms2));
        }
    }
    return 0;
}
