// Host-thread scatter of a DMA'd ring of victims into the pinned host table (writeback_pool.cpp).
#pragma once
#include <stdint.h>

namespace cebag {

struct WritebackJob {
    float* host_table;            // host address of the pinned table fp32[N, D]
    float* host_state;            // host address of the row-wise Adagrad state fp32[N], or null
    const float* ring;            // pinned ring: ring_rows rows of D floats, victims in ascending host-row order
    const float* ring_state;      // pinned fp32[ring_rows] or null
    const int32_t* rows;          // pinned int32[ring_rows]: host row of every ring entry
    int64_t ring_rows;            // rows that were DMA'd (an estimate made before E was known)
    int32_t dim;
    const int64_t* evicted;       // -> cebag_prepare_result.evicted of the call (pinned, device-written)
    const int64_t* status;        // -> cebag_prepare_result.status
};

// scatters min(evicted, ring_rows) rows; blocks until they are in the table
void run_writeback_job(const WritebackJob& job);

}  // namespace cebag
