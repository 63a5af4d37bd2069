// Host side of the DMA write-back: a small pool of threads that scatters rows from a pinned ring buffer (filled by a
// contiguous cudaMemcpyAsync D2H on the copy engine) into their rows of the pinned host table.
//
// Why the CPU: SM-issued PCIe reads and writes share one request path (a gather and a scatter kernel together reach
// 16 + 16 GB/s) and random rows of a 91 GB table cost one GPU-side address translation per 2 MB page (27-30 GB/s each
// way, serialised), whereas the copy engines move a contiguous buffer at 48-57 GB/s in parallel with the gather kernel
// (scripts/probes/pcie_swap_probe.cu, pcie_bigtable_probe.cu).  Scattering 160 k rows of 512 B with non-temporal
// AVX-512 stores takes 4 threads ~1.8 ms (scripts/probes/host_scatter_probe.cu) -- off the GPU's critical path.
// The job runs inside a cudaLaunchHostFunc callback (no CUDA calls here), so the event recorded after it on the
// stream means "the rows are in the table".
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include "writeback_pool.h"

namespace cebag {

namespace {

__attribute__((target("avx512f"))) void copy_row_nt512(char* dst, const char* src, size_t bytes) {
    size_t k = 0;
    for (; k + 64 <= bytes; k += 64)
        _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + k), _mm512_loadu_si512(reinterpret_cast<const void*>(src + k)));
    if (k < bytes) memcpy(dst + k, src + k, bytes - k);
}

struct Pool {
    std::mutex mu;
    std::condition_variable work_cv, done_cv;
    std::vector<std::thread> threads;
    const WritebackJob* job = nullptr;
    int64_t rows = 0;
    std::atomic<int64_t> next{0};
    int generation = 0, running = 0;
    bool use_nt = false;

    Pool() {
        const char* v = getenv("CEBAG_WB_THREADS");
        int n = (v && *v) ? atoi(v) : 4;
        if (n < 1) n = 1;
        if (n > 64) n = 64;
        use_nt = __builtin_cpu_supports("avx512f");
        for (int t = 0; t < n; ++t) threads.emplace_back([this]() { loop(); });
        for (auto& th : threads) th.detach();
    }

    void scatter_range(const WritebackJob& j, int64_t lo, int64_t hi) const {
        const size_t row_bytes = (size_t)j.dim * sizeof(float);
        const bool nt = use_nt && row_bytes % 64 == 0 && (reinterpret_cast<uintptr_t>(j.host_table) % 64 == 0);
        for (int64_t i = lo; i < hi; ++i) {
            const int64_t row = j.rows[i];
            char* dst = reinterpret_cast<char*>(j.host_table) + (size_t)row * row_bytes;
            const char* src = reinterpret_cast<const char*>(j.ring) + (size_t)i * row_bytes;
            if (nt) copy_row_nt512(dst, src, row_bytes);
            else memcpy(dst, src, row_bytes);
            if (j.host_state && j.ring_state) j.host_state[row] = j.ring_state[i];
        }
        if (nt) _mm_sfence();
    }

    void loop() {
        int seen = 0;
        for (;;) {
            const WritebackJob* j;
            int64_t n;
            {
                std::unique_lock<std::mutex> lock(mu);
                work_cv.wait(lock, [&]() { return generation != seen; });
                seen = generation;
                j = job;
                n = rows;
            }
            constexpr int64_t kGrain = 256;
            for (;;) {
                const int64_t lo = next.fetch_add(kGrain);
                if (lo >= n) break;
                scatter_range(*j, lo, lo + kGrain < n ? lo + kGrain : n);
            }
            {
                std::lock_guard<std::mutex> lock(mu);
                if (--running == 0) done_cv.notify_all();
            }
        }
    }

    void run(const WritebackJob& j, int64_t n) {
        const char* d = getenv("CEBAG_WB_DELAY_US");      // test hook: a slow write-back exposes ordering bugs
        const int delay_us = (d && *d) ? atoi(d) : 0;
        if (delay_us > 0) usleep(delay_us);
        const char* sk = getenv("CEBAG_WB_SKIP");         // probe hook: measure the DMA alone (tables end up wrong)
        if (sk && *sk == '1') return;
        if (n <= 0) return;
        std::unique_lock<std::mutex> lock(mu);
        job = &j;
        rows = n;
        next.store(0);
        running = (int)threads.size();
        ++generation;
        work_cv.notify_all();
        done_cv.wait(lock, [&]() { return running == 0; });
        job = nullptr;
    }
};

Pool& pool() {
    static Pool* p = new Pool();      // never destroyed: its threads outlive static destruction
    return *p;
}

}  // namespace

void run_writeback_job(const WritebackJob& job) {
    // the device wrote the record before the copies that precede this callback in stream order
    int64_t e = *reinterpret_cast<const volatile int64_t*>(job.evicted);
    const int64_t status = *reinterpret_cast<const volatile int64_t*>(job.status);
    if (status != 0) e = 0;
    int64_t n = e < job.ring_rows ? e : job.ring_rows;
    pool().run(job, n);
}

}  // namespace cebag
