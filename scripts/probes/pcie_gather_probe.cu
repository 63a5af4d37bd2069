// Probe (not product code): how fast can a kernel gather random 512 B rows from pinned host memory over PCIe,
// and scatter rows back, with different access patterns?  nvcc -arch=sm_100a -O3 -o pcie_probe pcie_gather_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// warp per row, ROWS rows in flight per warp (LDG.128 per lane)
template <int ROWS>
__global__ void gather_ldg(const float4* __restrict__ host, const int* __restrict__ rows, float4* __restrict__ dst, int m) {
    int lane = threadIdx.x & 31;
    long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long nw = ((long)gridDim.x * blockDim.x) >> 5;
    for (long j = warp * ROWS; j < m; j += nw * ROWS) {
        float4 v[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) if (j + r < m) v[r] = host[(long)rows[j + r] * 32 + lane];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) if (j + r < m) dst[(j + r) * 32 + lane] = v[r];
    }
}
// both directions: read host row -> device, write device row -> host row
__global__ void swap_ldg(float4* __restrict__ host, const int* __restrict__ rows_in, const int* __restrict__ rows_out,
                         float4* __restrict__ dev, int m) {
    int lane = threadIdx.x & 31;
    long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long nw = ((long)gridDim.x * blockDim.x) >> 5;
    for (long j = warp; j < m; j += nw) {
        float4 vin = host[(long)rows_in[j] * 32 + lane];
        float4 vout = dev[j * 32 + lane];
        host[(long)rows_out[j] * 32 + lane] = vout;
        dev[j * 32 + lane] = vin;
    }
}
// TMA-style bulk copy: one thread per row issues cp.async.bulk global->shared (512 B), then shared->global
__global__ void gather_bulk(const float4* __restrict__ host, const int* __restrict__ rows, float4* __restrict__ dst, int m,
                            int rows_per_cta) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    float4* buf = reinterpret_cast<float4*>(smem);
    unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&bar);
    for (long base = (long)blockIdx.x * rows_per_cta; base < m; base += (long)gridDim.x * rows_per_cta) {
        int cnt = (int)min((long)rows_per_cta, m - base);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr));
            asm volatile("fence.mbarrier_init.release.cluster;");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_addr), "r"(cnt * 512));
        }
        __syncthreads();
        for (int r = threadIdx.x; r < cnt; r += blockDim.x) {
            unsigned dsts = (unsigned)__cvta_generic_to_shared(buf + r * 32);
            const float4* src = host + (long)rows[base + r] * 32;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dsts), "l"(src), "r"(512), "r"(bar_addr) : "memory");
        }
        // wait
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar_addr), "r"(0));
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 32; i += blockDim.x) dst[base * 32 + i] = buf[i];
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(bar_addr));
        __syncthreads();
    }
}

int main(int argc, char** argv) {
    const long N = 16L << 20;          // 16 Mi rows x 512 B = 8 GiB pinned
    const int M = 160000;
    float4* host; CK(cudaHostAlloc(&host, N * 512, cudaHostAllocMapped | cudaHostAllocPortable));
    float4* hdev; CK(cudaHostGetDevicePointer(&hdev, host, 0));
    for (long i = 0; i < N * 32; i += 1024) host[i].x = (float)i;
    std::vector<int> r1(M), r2(M);
    srand(1);
    for (int i = 0; i < M; ++i) { r1[i] = (int)(((long)rand() * 65536 + rand()) % N); r2[i] = (int)(((long)rand() * 65536 + rand()) % N); }
    std::sort(r1.begin(), r1.end());
    int *d1, *d2; CK(cudaMalloc(&d1, M * 4)); CK(cudaMalloc(&d2, M * 4));
    CK(cudaMemcpy(d1, r1.data(), M * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d2, r2.data(), M * 4, cudaMemcpyHostToDevice));
    float4* dst; CK(cudaMalloc(&dst, (long)M * 512));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char* name, float ms, double dirs) {
        printf("%-44s %7.3f ms  %6.1f GB/s per direction\n", name, ms, M * 512.0 / ms / 1e6);
    };
#define TIME(name, launch) do { launch; CK(cudaDeviceSynchronize()); cudaEventRecord(e0); launch; cudaEventRecord(e1); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, e0, e1); report(name, ms, 1); } while (0)
    int grids[] = {37, 74, 148, 296, 592, 1184};
    for (int g : grids) {
        char nm[96];
        snprintf(nm, 96, "gather ldg 1 row/warp grid %d x128", g); TIME(nm, (gather_ldg<1><<<g, 128>>>(hdev, d1, dst, M)));
        snprintf(nm, 96, "gather ldg 4 rows/warp grid %d x128", g); TIME(nm, (gather_ldg<4><<<g, 128>>>(hdev, d1, dst, M)));
        snprintf(nm, 96, "swap (both dirs) grid %d x128", g); TIME(nm, (swap_ldg<<<g, 128>>>(hdev, d1, d2, dst, M)));
    }
    for (int g : {37, 74, 148, 296}) for (int rpc : {32, 128, 256}) {
        char nm[96];
        snprintf(nm, 96, "gather bulk (TMA 512B) grid %d rows/cta %d", g, rpc);
        CK(cudaFuncSetAttribute(gather_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, rpc * 512));
        TIME(nm, (gather_bulk<<<g, 128, rpc * 512>>>(hdev, d1, dst, M, rpc)));
    }
    // contiguous cudaMemcpy for reference
    cudaEventRecord(e0); CK(cudaMemcpyAsync(dst, host, (long)M * 512, cudaMemcpyHostToDevice)); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); report("cudaMemcpy H2D contiguous 82 MB", ms, 1);
    return 0;
}
