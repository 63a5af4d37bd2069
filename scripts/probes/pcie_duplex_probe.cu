// Probe (not product code): can the two PCIe directions run at full speed at once, and with which engines?
//   (1) zero-copy gather kernel (H2D)            alone
//   (2) zero-copy scatter kernel (D2H)           alone
//   (3) cudaMemcpyAsync D2H (copy engine), contiguous 82 MB   alone
//   (4) gather kernel || scatter kernel          (two streams)
//   (5) gather kernel || cudaMemcpyAsync D2H     (two streams)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_duplex pcie_duplex_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
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

int main() {
    const long N = 16L << 20;
    const int M = 160000;
    float4* host; CK(cudaHostAlloc(&host, N * 512, cudaHostAllocMapped | cudaHostAllocPortable));
    float4* hdev; CK(cudaHostGetDevicePointer(&hdev, host, 0));
    float4* stage; CK(cudaHostAlloc(&stage, (long)M * 512, cudaHostAllocPortable));
    std::vector<int> r1(M), r2(M);
    srand(1);
    for (int i = 0; i < M; ++i) { r1[i] = (int)(((long)rand() * 65536 + rand()) % N); r2[i] = (int)(((long)rand() * 65536 + rand()) % N); }
    std::sort(r1.begin(), r1.end());
    int *d1, *d2; CK(cudaMalloc(&d1, M * 4)); CK(cudaMalloc(&d2, M * 4));
    CK(cudaMemcpy(d1, r1.data(), M * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d2, r2.data(), M * 4, cudaMemcpyHostToDevice));
    float4 *a, *b; CK(cudaMalloc(&a, (long)M * 512)); CK(cudaMalloc(&b, (long)M * 512));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1, f0, f1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f0); cudaEventCreate(&f1);
    const int G = 56, T = 128;
    auto gbs = [&](float ms) { return M * 512.0 / ms / 1e6; };
    for (int rep = 0; rep < 2; ++rep) {
        float ms, ms2;
        cudaEventRecord(e0, s1); gather<<<G, T, 0, s1>>>(hdev, d1, a, M); cudaEventRecord(e1, s1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1); if (rep) printf("(1) gather kernel alone             %6.3f ms %5.1f GB/s\n", ms, gbs(ms));
        cudaEventRecord(e0, s1); scatter<<<G, T, 0, s1>>>(hdev, d2, b, M); cudaEventRecord(e1, s1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1); if (rep) printf("(2) scatter kernel alone            %6.3f ms %5.1f GB/s\n", ms, gbs(ms));
        cudaEventRecord(e0, s1); CK(cudaMemcpyAsync(stage, b, (long)M * 512, cudaMemcpyDeviceToHost, s1)); cudaEventRecord(e1, s1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1); if (rep) printf("(3) cudaMemcpyAsync D2H alone       %6.3f ms %5.1f GB/s\n", ms, gbs(ms));
        cudaEventRecord(e0, s1); cudaEventRecord(f0, s2);
        gather<<<G, T, 0, s1>>>(hdev, d1, a, M); scatter<<<G, T, 0, s2>>>(hdev, d2, b, M);
        cudaEventRecord(e1, s1); cudaEventRecord(f1, s2); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1); cudaEventElapsedTime(&ms2, f0, f1);
        if (rep) printf("(4) gather || scatter kernels       gather %6.3f ms %5.1f GB/s   scatter %6.3f ms %5.1f GB/s\n", ms, gbs(ms), ms2, gbs(ms2));
        cudaEventRecord(e0, s1); cudaEventRecord(f0, s2);
        gather<<<G, T, 0, s1>>>(hdev, d1, a, M); CK(cudaMemcpyAsync(stage, b, (long)M * 512, cudaMemcpyDeviceToHost, s2));
        cudaEventRecord(e1, s1); cudaEventRecord(f1, s2); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1); cudaEventElapsedTime(&ms2, f0, f1);
        if (rep) printf("(5) gather kernel || memcpy D2H     gather %6.3f ms %5.1f GB/s   memcpy  %6.3f ms %5.1f GB/s\n", ms, gbs(ms), ms2, gbs(ms2));
        cudaEventRecord(e0, s1); cudaEventRecord(f0, s2);
        CK(cudaMemcpyAsync(a, stage, (long)M * 512, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(stage, b, (long)M * 512, cudaMemcpyDeviceToHost, s2));
        cudaEventRecord(e1, s1); cudaEventRecord(f1, s2); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1); cudaEventElapsedTime(&ms2, f0, f1);
        if (rep) printf("(6) memcpy H2D || memcpy D2H        h2d    %6.3f ms %5.1f GB/s   d2h     %6.3f ms %5.1f GB/s\n", ms, gbs(ms), ms2, gbs(ms2));
    }
    return 0;
}
