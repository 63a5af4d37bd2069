/*
 * cebag.h -- C ABI of the B200-native cached embedding-bag hot path (libcebag_b200.so).
 *
 * Plain pointers and sizes only: no torch / C++ types cross this boundary.  Device pointers are raw
 * CUDA addresses, `stream` is a cudaStream_t passed as void*.  Every entry point returns a cebag_status
 * (0 = ok) and never throws; cebag_last_error() gives the message of the last failure on the calling thread.
 *
 * The reference (hpcaitech/CachedEmbedding) has no FFI for this path: its boundary is the Python nn.Module
 * surface of ColossalAI's cache_embedding package (pinned in /root/reference/README.md:37, absent from the
 * tree; behaviour restated in SURVEY.md Appendix A).  Each entry point below cites the reference call site /
 * upstream method it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 */
#ifndef CEBAG_H_
#define CEBAG_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CEBAG_ABI_VERSION 8

#if defined(__GNUC__)
#define CEBAG_API __attribute__((visibility("default")))
#else
#define CEBAG_API
#endif

typedef enum cebag_status {
    CEBAG_OK = 0,
    CEBAG_ERR_INVALID = 1,   /* bad argument (null pointer, bad size, unsupported mode)          */
    CEBAG_ERR_CUDA = 2,      /* a CUDA runtime call or kernel launch failed                       */
    CEBAG_ERR_CAPACITY = 3,  /* unique rows of one prepare_ids call exceed the cache (A.3 assert) */
    CEBAG_ERR_INDEX = 4      /* an id outside [0, num_rows) was seen                               */
} cebag_status;

enum { CEBAG_EVICT_LFU = 1, CEBAG_EVICT_DATASET = 2 };        /* EvictionStrategy (recsys/models/dlrm.py:66,80) */
enum { CEBAG_MODE_SUM = 0, CEBAG_MODE_MEAN = 1 };              /* F.embedding_bag mode                           */
enum { CEBAG_OPT_SGD = 0, CEBAG_OPT_ROWWISE_ADAGRAD = 1 };
#define CEBAG_MAX_PEERS 8
enum { CEBAG_LAYOUT_BAG_MAJOR = 0,     /* out[g, :]                       -- what F.embedding_bag returns            */
       CEBAG_LAYOUT_SAMPLE_MAJOR = 1,  /* out[(g % B) * F + g / B, :]     -- the (B, F, D) view the DLRM shape hooks
                                          build (recsys/models/dlrm.py:26-30), written directly by the kernel   */
       CEBAG_LAYOUT_EXCHANGE = 2       /* table-wise sharding with the all-to-all FUSED into the kernel: bag (f, b) of
                                          this rank's f-th table lives in the buffer of the rank that owns sample b
                                          (cebag_exchange), reached over NVLink peer memory                     */ };

#define CEBAG_FREQ_EMPTY INT64_MAX     /* LFU counter of an empty slot (upstream: sys.maxsize)                   */

/*
 * State of one cached table.  All arrays are owned by the caller (the Python module allocates them as torch
 * tensors); the library only reads/writes through these pointers.  Mirrors CachedParamMgr's buffers (A.1):
 *   weight -> host_table, cuda_cached_weight -> cache, idx_map, cached_idx_map -> slot2row,
 *   inverted_cached_idx -> row2slot, freq_cnter -> freq.
 * B200-first differences: the id maps are int32 (half the HBM of the reference's int64 maps, which were 76 % of
 * its footprint); idx_map may be NULL (identity); a per-slot window stamp, a per-slot hit flag
 * and a per-row miss bitmap replace the reference's sort-based unique/isin; the free-slot count and the window stamp
 * live in DEVICE memory (dev_state), so that prepare_ids never has to wait for the GPU.
 */
enum { CEBAG_STATE_AVAIL = 0,   /* free slots (upstream _cuda_available_row_num)                        */
       CEBAG_STATE_EPOCH = 1,   /* window stamp of the last successful prepare_ids                      */
       CEBAG_STATE_CALLS = 2,   /* prepare_ids calls that completed on the device, successful or not    */
       CEBAG_STATE_MAXFREQ = 3, /* upper bound of the LFU counters of occupied slots (victim selection skips the
                                   radix passes above its highest byte)                                  */
       CEBAG_STATE_WORDS = 8 };

typedef struct cebag_table {
    int64_t   num_rows;       /* N: rows of the host table                                                */
    int32_t   dim;            /* D: floats per row                                                        */
    int32_t   cache_rows;     /* C: slots in HBM                                                          */
    int32_t   strategy;       /* CEBAG_EVICT_*                                                            */
    int32_t   protect_windows;/* slots stamped by the last `protect_windows` prepare_ids calls are never evicted:
                                 1 = the reference's rule (only the current call's rows, A.4);
                                 2 = look-ahead overlap: the previous window may still be computing on them    */
    float*    host_table;     /* fp32[N, D]  pinned host memory, device-visible address                   */
    float*    host_state;     /* fp32[N]     row-wise Adagrad state in pinned host memory, or NULL        */
    float*    cache;          /* fp32[C, D]  HBM                                                          */
    float*    cache_state;    /* fp32[C]     HBM, or NULL                                                 */
    const int32_t* idx_map;   /* int32[N] id -> row, or NULL for identity                                 */
    int32_t*  row2slot;       /* int32[N]   < 0 = not resident (-1; values <= -2 are transient markers of
                                 rows being written back)                                                 */
    int32_t*  slot2row;       /* int32[C]   -1 = empty                                                    */
    int64_t*  freq;           /* int64[C]   LFU counters (CEBAG_FREQ_EMPTY = empty); NULL for DATASET     */
    int32_t*  slot_epoch;     /* int32[C]   stamp of the last prepare_ids window that used the slot       */
    uint32_t* miss_bitmap;    /* uint32[ceil(N/32)] all-zero between calls                                */
    uint8_t*  hit_flags;      /* uint8[16 * ceil(C/16)] all-zero between calls: slots hit by the call     */
    int64_t*  dev_state;      /* int64[CEBAG_STATE_WORDS] HBM; [AVAIL] = C and the rest 0 for an empty cache */
} cebag_table;

/* Scratch and stream plumbing of one prepare_ids / flush call; sizes from cebag_prepare_workspace_bytes().
 * Everything a call enqueues reads its row lists from `device`: the buffer must stay untouched until the work has
 * run (until writeback_done_event when a copy stream is used). */
typedef struct cebag_workspace {
    void*  device;            /* device scratch                                                           */
    size_t device_bytes;
    void*  pinned;            /* >= 256 bytes of pinned, device-mapped host memory (result read-back of the
                                 synchronous entry points)                                                */
    void*  copy_stream;       /* optional cudaStream_t: the PCIe row traffic of prepare_ids (fill of the missed rows,
                                 write-back of the victims) is enqueued here, after the maps are committed on
                                 `stream`, instead of on `stream`                                             */
    void*  copy_done_event;   /* optional cudaEvent_t recorded on copy_stream once the missed rows are in HBM:
                                 what the forward of the window has to wait for                              */
    void*  writeback_done_event; /* optional cudaEvent_t recorded on copy_stream after the victims have reached the
                                 host table (and `device` / `stage` may be reused)                           */
    void*  victims_ready_event;  /* optional cudaEvent_t: `stream` waits for it before the victims' rows are read
                                 (their last optimizer update may belong to an earlier, still running window)   */
    float* stage;             /* optional fp32[stage_rows, D] HBM (copy_stream only): victims are parked here so that
                                 the fill does not wait for their write-back over PCIe                        */
    float* stage_state;       /* fp32[stage_rows] or NULL (row-wise Adagrad state of the parked victims)   */
    int64_t stage_rows;
    /* DMA write-back (needs copy_stream + stage): the first dma_rows parked victims leave through a copy engine -- one
     * contiguous cudaMemcpyAsync into a pinned ring -- and host threads scatter them into the table; the rest of the
     * parked victims is written by the zero-copy kernel as before.  dma_rows is the caller's ESTIMATE of E (E itself
     * only exists on the device): the host side scatters min(E, dma_rows) rows.  While a write-back is in flight the
     * victim's row2slot entry is the marker -2 - (its index in `stage`): a later call that re-admits the row fills it
     * from that staging buffer (prev_stage) instead of the host table; `retire_device` names the workspace of an
     * earlier call whose markers are cleared once its write-back has completed (retire_wait_event). */
    void*  dma_stream;        /* cudaStream_t of the D2H copies and of the host scatter                    */
    void*  dma_done_event;    /* cudaEvent_t recorded on dma_stream once those rows are in the host table   */
    void*  dma_wait_event;    /* optional cudaEvent_t dma_stream waits for first (zero-copy write-backs of the
                                 previous call, which may target the same host rows)                        */
    float* dma_ring;          /* pinned host fp32[dma_rows, D]                                             */
    float* dma_ring_state;    /* pinned host fp32[dma_rows] or NULL                                        */
    int32_t* dma_ring_rows;   /* pinned host int32[dma_rows]: host row of every ring entry                 */
    int64_t dma_rows;
    float* host_table_hostptr;/* HOST addresses of cebag_table.host_table / host_state (for the host threads) */
    float* host_state_hostptr;
    const float* prev_stage;  /* staging buffer the markers currently in row2slot point into, or NULL      */
    const float* prev_stage_state;
    const void* retire_device;/* `device` of the earlier call whose markers are to be cleared, or NULL     */
    void*  retire_wait_event; /* cudaEvent_t `stream` waits for before clearing them                       */
} cebag_workspace;

/* What one prepare_ids call did (upstream: num_hits_history / num_miss_history / num_write_back_history,
 * _cache_miss, _total_cache, _cpu_to_cuda_numel, _cuda_to_cpu_numel). */
typedef struct cebag_prepare_stats {
    int64_t unique_hits;      /* unique rows of the call already resident                                 */
    int64_t unique_misses;    /* unique rows brought in (M)                                               */
    int64_t evicted;          /* rows written back to the host table (E)                                  */
    int64_t miss_lookups;     /* ids (with multiplicity) whose row was not resident                       */
    int64_t total_lookups;    /* n                                                                        */
} cebag_prepare_stats;

/* Result record of cebag_prepare_ids_async: written by the DEVICE into pinned, device-mapped host memory when the
 * call's map work has run; `status` is CEBAG_PREPARE_PENDING until then. */
#define CEBAG_PREPARE_PENDING (-1)
typedef struct cebag_prepare_result {
    int64_t status;           /* CEBAG_OK, CEBAG_ERR_CAPACITY, CEBAG_ERR_INDEX or CEBAG_PREPARE_PENDING   */
    int64_t unique_hits, unique_misses, evicted, miss_lookups, total_lookups;
    int64_t evictable;        /* occupied slots outside the protected windows (protect_windows > 1)       */
    int64_t avail_after;      /* free slots after the call                                                */
} cebag_prepare_result;

CEBAG_API int         cebag_abi_version(void);
CEBAG_API const char* cebag_last_error(void);

/* ---- launch accounting and per-kernel timers (upstream: CachedParamMgr._elapsed_dict / print_comm_stats) -------
 * cebag_launch_count: kernels launched by this library since it was loaded.
 * With profiling enabled every group of launches is bracketed by CUDA events on its stream; cebag_profile_collect
 * waits for them, fills total_ms[k] / launches[k] for k < cebag_profile_num_kernels() and resets the timers.
 * Entry k covers the kernels named cebag_profile_kernel_name(k) ("radix_sort" = all kernels of one sort). */
CEBAG_API int64_t     cebag_launch_count(void);
CEBAG_API int         cebag_profile_enable(int on);
CEBAG_API int         cebag_profile_num_kernels(void);
CEBAG_API const char* cebag_profile_kernel_name(int k);
CEBAG_API int         cebag_profile_collect(double* total_ms, int64_t* launches);

/* ---- pinned host memory for the table (upstream: weight.pin_memory(), A.1) --------------------------------- */
CEBAG_API int cebag_host_alloc(void** out_ptr, size_t bytes);                 /* cudaHostAlloc(portable|mapped)          */
CEBAG_API int cebag_host_free(void* ptr);
CEBAG_API int cebag_host_register(void* ptr, size_t bytes);                   /* pin an existing allocation in place     */
CEBAG_API int cebag_host_unregister(void* ptr);
CEBAG_API int cebag_host_device_pointer(void* host_ptr, void** out_dev_ptr);  /* device-visible alias of a pinned ptr    */

/* Device memory that other processes of the node can map (plain cudaMalloc + CUDA IPC), for cebag_exchange.peer. */
CEBAG_API int cebag_device_alloc(void** out_ptr, size_t bytes);
CEBAG_API int cebag_device_free(void* ptr);
CEBAG_API int cebag_ipc_export(void* ptr, unsigned char handle[64]);
CEBAG_API int cebag_ipc_import(const unsigned char handle[64], void** out_ptr);
CEBAG_API int cebag_ipc_close(void* ptr);

/* Fill fp32[count] (device or pinned host memory) with U(lo, hi) from a counter-based generator; the value of
 * element i depends only on (seed, i).  Replaces _weight_alloc's uniform_(-1/N, 1/N) (A.2) for tables that are
 * too large to initialise from one host thread. */
CEBAG_API int cebag_fill_uniform(float* dst, int64_t count, float lo, float hi, uint64_t seed, void* stream);

/* id -> frequency histogram on the GPU: freq[id] += 1 for every id (int64[n], device); *bad_flag (device int32) is set
 * when an id falls outside [0, num_rows).  Replaces the np.bincount counters of recsys/datasets/feature_counter.py:21-29
 * whose output (ids_freq_mapping, int64[N]) feeds CachedParamMgr.reorder (A.1). */
CEBAG_API int cebag_id_histogram(const int64_t* ids, int64_t n, int64_t* freq, int64_t num_rows, int32_t* bad_flag,
                       void* stream);

/* ---- cache manager (CachedParamMgr) ---------------------------------------------------------------------------- */
CEBAG_API size_t cebag_prepare_workspace_bytes(const cebag_table* t, int64_t n_ids);

/* CachedParamMgr.prepare_ids(ids) (A.3 + A.4; called at recsys/dlrm_main.py:259 and inside forward when cache_op).
 * ids: int64[n] on the device.  slot_ids_out: int64[n] on the device, slot of every id, same order.
 * Performs unique / miss detection / victim selection / write-back of victims to the host table / fill of missed
 * rows / map + LFU counter update.
 *
 * cebag_prepare_ids_async only ENQUEUES work and never waits for the GPU: every size the kernels need (misses,
 * evictions, free slots) stays in device memory, grids are upper bounds, and the state-changing kernels are
 * conditional on a device-side verdict -- an id out of range or a window that does not fit the cache (A.3 assert)
 * leaves the table exactly as it was (slot_ids_out is then unspecified).  `result` (pinned, device-mapped) is set to
 * PENDING by the call and filled in stream order; cebag_prepare_result_status() turns it into a status + message.
 * cebag_prepare_ids = the same + one wait for `stream` at the end, returning the verdict and `stats`. */
CEBAG_API int cebag_prepare_ids_async(const cebag_table* t, const int64_t* ids, int64_t n, int64_t* slot_ids_out,
                            const cebag_workspace* ws, cebag_prepare_result* result, void* stream);
CEBAG_API int cebag_prepare_result_status(const cebag_table* t, const cebag_prepare_result* result,
                            cebag_prepare_stats* stats_out);
CEBAG_API int cebag_prepare_ids(const cebag_table* t, const int64_t* ids, int64_t n, int64_t* slot_ids_out,
                      const cebag_workspace* ws, cebag_prepare_stats* stats, void* stream);

/* CachedParamMgr.flush() (A.1): write every resident row (and state) back to the host table, empty the maps.
 * Returns the number of rows written in *rows_written.  Synchronises `stream`. */
CEBAG_API int cebag_flush(const cebag_table* t, const cebag_workspace* ws, int64_t* rows_written, void* stream);

/* Warm-up preload of CachedParamMgr.reorder() (A.1 step 2): rows[k] (int32, device) go to slots 0..k-1;
 * freq_init (int64[k], device) seeds the LFU counters, NULL means 0.  The cache must be empty. */
CEBAG_API int cebag_preload(const cebag_table* t, const int32_t* rows, const int64_t* freq_init, int64_t k, void* stream);

/* Single-row legacy helpers of upstream test_cachemgr (B.1): _admit(row) into `slot`, _evict of `slot`. */
CEBAG_API int cebag_admit_row(const cebag_table* t, int64_t row, int64_t slot, void* stream);
CEBAG_API int cebag_evict_slot(const cebag_table* t, int64_t slot, void* stream);

/* Free slots right now (waits for `stream`; upstream cuda_available_row_num). */
CEBAG_API int cebag_available_rows(const cebag_table* t, int64_t* avail_out, void* stream);

/* Peer buffers of the fused pooled-embedding exchange (replaces dual_all_to_all_tablewise, A.6 / SURVEY K15).
 * Rank j owns samples [start_j, start_j + B_j) of the global batch B (torch.tensor_split rule: the first B % world
 * ranks hold one more) and a buffer fp32[B_j, total_features, D] with features in rank-major order; peer[j] is that
 * buffer mapped into THIS process (CUDA IPC over NVLink; peer[own rank] is the local buffer).  The forward stores
 * every pooled row of this rank's tables straight into the owner's buffer; the backward loads every gradient row
 * straight from it.  The caller orders the kernels of different ranks (a barrier after the forward's stores and
 * before the backward's loads). */
typedef struct cebag_exchange {
    int32_t world;                   /* ranks, <= CEBAG_MAX_PEERS                                              */
    int32_t feature_offset;          /* position of this rank's first table in the rank-major feature order    */
    int32_t total_features;          /* F over all ranks                                                       */
    int32_t reserved0;
    float*  peer[CEBAG_MAX_PEERS];
} cebag_exchange;

/* Stream-ordered barrier of the ranks of a fused exchange, over peer memory: flags->peer[j] is rank j's array of
 * CEBAG_MAX_PEERS uint32 flags (zero-initialised, mapped into this process); *seq_counter (device uint32, zero at
 * start, private to this rank) counts the barriers and is advanced by the kernel itself, so the call can be captured
 * in a CUDA graph and replayed.
 * Everything enqueued on `stream` before the call, on every rank, is complete and visible before anything enqueued
 * after it starts.  *failed_flag (device int32) is set if a peer does not arrive within a few seconds. */
CEBAG_API int cebag_peer_barrier(const cebag_exchange* flags, int32_t rank, uint32_t* seq_counter, int32_t* failed_flag,
                       void* stream);

/* ---- embedding bag over the slot cache (F.embedding_bag on cuda_cached_weight, A.2) --------------------------- */
typedef struct cebag_bag_args {
    const float*   cache;          /* fp32[C, D]                                                            */
    int32_t        cache_rows;     /* C                                                                     */
    int32_t        dim;            /* D                                                                     */
    const int64_t* slot_ids;       /* int64[n]                                                              */
    int64_t        n;              /* lookups                                                               */
    const void*    offsets;        /* int32 or int64 [num_bags] or [num_bags + 1]                           */
    int32_t        offsets_are_64; /* 1: int64 offsets, 0: int32                                            */
    int32_t        include_last_offset;
    int64_t        num_bags;       /* G                                                                     */
    const float*   per_sample_weights; /* fp32[n] or NULL                                                   */
    int32_t        mode;           /* CEBAG_MODE_*                                                          */
    int64_t        padding_idx;    /* slot id to skip, or -1                                                */
    int32_t        layout;         /* CEBAG_LAYOUT_* of out / grad_out                                      */
    int64_t        layout_batch;   /* B for CEBAG_LAYOUT_SAMPLE_MAJOR / _EXCHANGE (num_bags == F_local * B)  */
    const cebag_exchange* exchange;/* CEBAG_LAYOUT_EXCHANGE only: where out (forward) / grad_out (backward) live;
                                      the out / grad_out pointer arguments are then ignored                  */
    /* backward with workspace_has_plan == 2: this batch's segment of a WINDOW plan (cebag_bag_backward_plan_window) */
    const uint32_t* plan_keys;     /* uint32[n] sorted keys of the batch: (batch index << bits) | slot         */
    const uint32_t* plan_vals;     /* uint32[n] bag of every sorted lookup                                     */
    uint32_t       plan_key_mask;  /* slot = key & plan_key_mask                                               */
    uint32_t       reserved1;
} cebag_bag_args;

/* forward: out fp32[G, D] (or sample-major).  Replaces F.embedding_bag (recsys/models/dlrm.py:99-110). */
CEBAG_API int cebag_bag_forward(const cebag_bag_args* a, float* out, void* stream);

CEBAG_API size_t cebag_backward_workspace_bytes(const cebag_bag_args* a);

/* backward fused with the optimizer step on the cached rows: replaces _embedding_bag_sparse_backward +
 * coalesce + torch.optim.SGD's sparse branch (recsys/dlrm_main.py:274-279; A.8).  cache is updated in place:
 *   SGD:             W[s] -= lr * sum_i w_i * grad_out[bag(i)]
 *   row-wise Adagrad: g = that sum;  state[s] += mean(g^2);  W[s] -= lr * g / (sqrt(state[s]) + eps)
 * Deterministic: lookups are radix-sorted by slot and reduced in index order; no float atomics. */
CEBAG_API int cebag_bag_backward_fused(const cebag_bag_args* a, const float* grad_out, float* cache_rw, float* cache_state,
                             int32_t optimizer, float lr, float eps, void* workspace, size_t workspace_bytes,
                             int32_t workspace_has_plan, void* stream);

/* The integer half of the fused backward (lookup -> bag map, radix sort of the lookups by slot) for one batch, left in
 * `workspace` (cebag_backward_workspace_bytes).  It depends only on slot_ids / offsets, not on the gradient, so a
 * look-ahead driver runs it on its side stream right after prepare_ids; cebag_bag_backward_fused called with the same
 * arguments, the same workspace and workspace_has_plan = 1 then starts at the segment-reduce kernels.
 * Only for mode sum without per-sample weights (returns CEBAG_ERR_INVALID otherwise). */
CEBAG_API int cebag_bag_backward_plan(const cebag_bag_args* a, void* workspace, size_t workspace_bytes, void* stream);

/* The same for a whole look-ahead window in ONE radix sort: the lookups of `num_batches` batches (an array of
 * cebag_bag_args, mode sum without per-sample weights) are sorted by (batch, slot); batch j's segment is returned in
 * keys_out[j] / vals_out[j] (device pointers into window_workspace, host arrays of num_batches entries) and *mask_out.
 * cebag_bag_backward_fused with plan_keys / plan_vals / plan_key_mask set to them, workspace_has_plan = 2 and a scratch
 * workspace of cebag_backward_workspace_bytes(a) then starts at the segment-reduce kernels.  One sort over 13.6 M pairs
 * costs a quarter of eight sorts over 1.7 M (the small sorts are latency-bound). */
CEBAG_API size_t cebag_backward_window_plan_bytes(int64_t total_lookups);
CEBAG_API int cebag_bag_backward_plan_window(const cebag_bag_args* batches, int32_t num_batches, void* window_workspace,
                                   size_t workspace_bytes, const uint32_t** keys_out, const uint32_t** vals_out,
                                   uint32_t* mask_out, void* stream);

/* backward, compatibility forms for an external torch optimizer:
 *   coo:   values fp32[n, D] with values[i] = w_i * grad_out[bag(i)] (indices are slot_ids) -- the COO grad that
 *          sparse=True produces;  dense: grad fp32[C, D] fully written (zeros + segment sums) for sparse=False. */
CEBAG_API int cebag_bag_backward_coo(const cebag_bag_args* a, const float* grad_out, float* values, void* stream);
CEBAG_API int cebag_bag_backward_dense(const cebag_bag_args* a, const float* grad_out, float* grad_cache, void* workspace,
                             size_t workspace_bytes, void* stream);

/* grad w.r.t. per_sample_weights (mode sum): gw[i] = <grad_out[bag(i)], cache[slot_i]>. */
CEBAG_API int cebag_bag_backward_weights(const cebag_bag_args* a, const float* grad_out, float* grad_weights, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CEBAG_H_ */
