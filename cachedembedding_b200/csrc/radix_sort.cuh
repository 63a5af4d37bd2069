// LSD radix sort of (slot id, lookup position) pairs (radix_sort.cu).
#pragma once
#include "common.cuh"

namespace cebag {

size_t radix_sort_workspace_bytes(int64_t n);

// Sorts lookups 0..n-1 by the low `key_bits` bits of slot_ids[i], stable.  The value carried with lookup i is
// init_vals[i], or i itself when init_vals is null.  On return *keys_sorted / *vals_sorted point into the
// workspace: keys_sorted[j] is the j-th smallest slot, vals_sorted[j] the value of that lookup.
int radix_sort_slots(const int64_t* slot_ids, int64_t n, int key_bits, void* workspace, size_t workspace_bytes,
                     const uint32_t* init_vals, const uint32_t** keys_sorted, const uint32_t** vals_sorted,
                     cudaStream_t stream);

// The same for pairs whose 32-bit keys the caller has already written: radix_sort_input_buffers gives the arrays to
// fill (they live inside the workspace), radix_sort_u32 sorts them on the low `key_bits` bits.
void radix_sort_input_buffers(int64_t n, void* workspace, uint32_t** keys_in, uint32_t** vals_in);
int radix_sort_u32(int64_t n, int key_bits, void* workspace, size_t workspace_bytes, const uint32_t** keys_sorted,
                   const uint32_t** vals_sorted, cudaStream_t stream);

// Where radix_sort_slots leaves its result inside `workspace` (for callers that sort now and consume later).
void radix_sort_result(int64_t n, int key_bits, void* workspace, const uint32_t** keys_sorted,
                       const uint32_t** vals_sorted);

}  // namespace cebag
