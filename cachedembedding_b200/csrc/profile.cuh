// Launch accounting and per-kernel CUDA-event timers (the native counterpart of the reference cache manager's
// named timers, SURVEY.md section 5 "Tracing / profiling").
#pragma once
#include "common.cuh"

namespace cebag {

enum KernelId : int {
    kKernForward = 0, kKernBagOf, kKernSort, kKernBwdPhase1, kKernBwdPhase2, kKernBwdCoo, kKernBwdWeights,
    kKernProbe, kKernBitmapRank, kKernSelect, kKernFreeSlots, kKernVictimRank, kKernPark, kKernFillRows, kKernWriteBack,
    kKernFixup, kKernLfuCount, kKernFlush, kKernMoveRows, kKernFill, kKernIdHistogram, kKernCount
};

void count_launches(int n);

// RAII: counts `launches` kernel launches and, when profiling is on, brackets them with events on `stream`.
class KernelScope {
public:
    KernelScope(int id, cudaStream_t stream, int launches = 1);
    ~KernelScope();
    KernelScope(const KernelScope&) = delete;
    KernelScope& operator=(const KernelScope&) = delete;
private:
    int id_;
    cudaStream_t stream_;
    cudaEvent_t stop_;
    bool timed_;
};

}  // namespace cebag
