#include <atomic>
#include <mutex>
#include <vector>
#include "profile.cuh"

namespace cebag {

namespace {

const char* const kNames[kKernCount] = {
    "bag_forward", "bag_of", "radix_sort", "bag_backward_phase1", "bag_backward_phase2", "bag_backward_coo",
    "bag_backward_weights", "probe", "bitmap_rank", "victim_select", "free_slots", "victim_rank", "park_victims",
    "fill_rows", "write_back", "fixup", "lfu_count", "flush", "move_rows", "fill_uniform", "id_histogram"};

std::atomic<long long> g_launches{0};
std::atomic<int> g_enabled{0};
std::mutex g_mu;
struct Pair { cudaEvent_t start, stop; };
std::vector<Pair> g_pending[kKernCount];
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

}  // namespace

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

KernelScope::KernelScope(int id, cudaStream_t stream, int launches)
    : id_(id), stream_(stream), stop_(nullptr), timed_(false) {
    g_launches.fetch_add(launches, std::memory_order_relaxed);
    if (g_enabled.load(std::memory_order_relaxed)) {
        std::lock_guard<std::mutex> lock(g_mu);
        Pair p{get_event(), get_event()};
        if (p.start && p.stop) {
            cudaEventRecord(p.start, stream_);
            g_pending[id_].push_back(p);
            stop_ = p.stop;
            timed_ = true;
        }
    }
}

KernelScope::~KernelScope() {
    if (timed_) cudaEventRecord(stop_, stream_);
}

}  // namespace cebag

using namespace cebag;

extern "C" int64_t cebag_launch_count(void) { return g_launches.load(); }

extern "C" int cebag_profile_enable(int on) {
    g_enabled.store(on ? 1 : 0);
    return CEBAG_OK;
}

extern "C" int cebag_profile_num_kernels(void) { return kKernCount; }

extern "C" const char* cebag_profile_kernel_name(int k) { return (k >= 0 && k < kKernCount) ? kNames[k] : ""; }

extern "C" int cebag_profile_collect(double* total_ms, int64_t* launches) {
    CEBAG_REQUIRE(total_ms != nullptr && launches != nullptr, "profile_collect arguments");
    std::lock_guard<std::mutex> lock(g_mu);
    for (int k = 0; k < kKernCount; ++k) {
        total_ms[k] = 0.0;
        launches[k] = 0;
        for (const Pair& p : g_pending[k]) {
            CEBAG_CUDA_CHECK(cudaEventSynchronize(p.stop));
            float ms = 0.f;
            CEBAG_CUDA_CHECK(cudaEventElapsedTime(&ms, p.start, p.stop));
            total_ms[k] += ms;
            launches[k] += 1;
            g_pool.push_back(p.start);
            g_pool.push_back(p.stop);
        }
        g_pending[k].clear();
    }
    return CEBAG_OK;
}
