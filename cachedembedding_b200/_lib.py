"""ctypes binding of libcebag_b200.so (C ABI declared in include/cebag.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing, or a call is made
without a CUDA device, this module raises -- it never routes around the CUDA path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcebag_b200.so")

ABI_VERSION = 8

# cebag_status
OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_INDEX = 0, 1, 2, 3, 4
PREPARE_PENDING = -1
STATE_AVAIL, STATE_EPOCH, STATE_CALLS, STATE_MAXFREQ, STATE_WORDS = 0, 1, 2, 3, 8
EVICT_LFU, EVICT_DATASET = 1, 2
MODE_SUM, MODE_MEAN = 0, 1
OPT_SGD, OPT_ROWWISE_ADAGRAD = 0, 1
LAYOUT_BAG_MAJOR, LAYOUT_SAMPLE_MAJOR, LAYOUT_EXCHANGE = 0, 1, 2
MAX_PEERS = 8
FREQ_EMPTY = 2**63 - 1

# every symbol include/cebag.h declares (tests check the library exports exactly these)
EXPORTS = (
    "cebag_abi_version", "cebag_last_error",
    "cebag_launch_count", "cebag_profile_enable", "cebag_profile_num_kernels", "cebag_profile_kernel_name",
    "cebag_profile_collect",
    "cebag_host_alloc", "cebag_host_free", "cebag_host_register", "cebag_host_unregister",
    "cebag_host_device_pointer", "cebag_fill_uniform", "cebag_id_histogram",
    "cebag_device_alloc", "cebag_device_free", "cebag_ipc_export", "cebag_ipc_import", "cebag_ipc_close",
    "cebag_peer_barrier",
    "cebag_prepare_workspace_bytes", "cebag_prepare_ids", "cebag_prepare_ids_async", "cebag_prepare_result_status",
    "cebag_flush", "cebag_preload", "cebag_admit_row", "cebag_evict_slot", "cebag_available_rows",
    "cebag_bag_forward", "cebag_backward_workspace_bytes", "cebag_bag_backward_fused", "cebag_bag_backward_plan",
    "cebag_backward_window_plan_bytes", "cebag_bag_backward_plan_window",
    "cebag_bag_backward_coo", "cebag_bag_backward_dense", "cebag_bag_backward_weights",
)


class CebagError(RuntimeError):
    """A libcebag_b200 call failed."""


class CacheCapacityError(AssertionError, CebagError):
    """Unique rows of one prepare_ids call exceed the cache (the reference raises this as an `assert`)."""


class Table(Structure):
    _fields_ = [
        ("num_rows", c_int64), ("dim", c_int32), ("cache_rows", c_int32), ("strategy", c_int32),
        ("protect_windows", c_int32),
        ("host_table", c_void_p), ("host_state", c_void_p), ("cache", c_void_p), ("cache_state", c_void_p),
        ("idx_map", c_void_p), ("row2slot", c_void_p), ("slot2row", c_void_p), ("freq", c_void_p),
        ("slot_epoch", c_void_p), ("miss_bitmap", c_void_p), ("hit_flags", c_void_p), ("dev_state", c_void_p),
    ]


class Workspace(Structure):
    _fields_ = [("device", c_void_p), ("device_bytes", c_size_t), ("pinned", c_void_p),
                ("copy_stream", c_void_p), ("copy_done_event", c_void_p), ("writeback_done_event", c_void_p),
                ("victims_ready_event", c_void_p), ("stage", c_void_p), ("stage_state", c_void_p),
                ("stage_rows", c_int64),
                ("dma_stream", c_void_p), ("dma_done_event", c_void_p), ("dma_wait_event", c_void_p),
                ("dma_ring", c_void_p), ("dma_ring_state", c_void_p), ("dma_ring_rows", c_void_p),
                ("dma_rows", c_int64), ("host_table_hostptr", c_void_p), ("host_state_hostptr", c_void_p),
                ("prev_stage", c_void_p), ("prev_stage_state", c_void_p), ("retire_device", c_void_p),
                ("retire_wait_event", c_void_p)]


class PrepareStats(Structure):
    _fields_ = [("unique_hits", c_int64), ("unique_misses", c_int64), ("evicted", c_int64),
                ("miss_lookups", c_int64), ("total_lookups", c_int64)]


class PrepareResult(Structure):
    """Device-written record of one cebag_prepare_ids_async call (pinned, device-mapped host memory)."""
    _fields_ = [("status", c_int64), ("unique_hits", c_int64), ("unique_misses", c_int64), ("evicted", c_int64),
                ("miss_lookups", c_int64), ("total_lookups", c_int64), ("evictable", c_int64),
                ("avail_after", c_int64)]


class Exchange(Structure):
    _fields_ = [("world", c_int32), ("feature_offset", c_int32), ("total_features", c_int32), ("reserved0", c_int32),
                ("peer", c_void_p * 8)]


class BagArgs(Structure):
    _fields_ = [
        ("cache", c_void_p), ("cache_rows", c_int32), ("dim", c_int32),
        ("slot_ids", c_void_p), ("n", c_int64),
        ("offsets", c_void_p), ("offsets_are_64", c_int32), ("include_last_offset", c_int32),
        ("num_bags", c_int64),
        ("per_sample_weights", c_void_p),
        ("mode", c_int32),
        ("padding_idx", c_int64),
        ("layout", c_int32),
        ("layout_batch", c_int64),
        ("exchange", POINTER(Exchange)),
        ("plan_keys", c_void_p), ("plan_vals", c_void_p), ("plan_key_mask", ctypes.c_uint32), ("reserved1", ctypes.c_uint32),
    ]


_lib = None


def _declare(lib):
    lib.cebag_abi_version.restype = c_int
    lib.cebag_last_error.restype = c_char_p
    lib.cebag_launch_count.restype = c_int64
    lib.cebag_profile_enable.argtypes = [c_int]
    lib.cebag_profile_num_kernels.restype = c_int
    lib.cebag_profile_kernel_name.argtypes = [c_int]
    lib.cebag_profile_kernel_name.restype = c_char_p
    lib.cebag_profile_collect.argtypes = [POINTER(ctypes.c_double), POINTER(c_int64)]
    lib.cebag_host_alloc.argtypes = [POINTER(c_void_p), c_size_t]
    lib.cebag_host_free.argtypes = [c_void_p]
    lib.cebag_host_register.argtypes = [c_void_p, c_size_t]
    lib.cebag_host_unregister.argtypes = [c_void_p]
    lib.cebag_host_device_pointer.argtypes = [c_void_p, POINTER(c_void_p)]
    lib.cebag_fill_uniform.argtypes = [c_void_p, c_int64, c_float, c_float, c_uint64, c_void_p]
    lib.cebag_id_histogram.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]
    lib.cebag_device_alloc.argtypes = [POINTER(c_void_p), c_size_t]
    lib.cebag_device_free.argtypes = [c_void_p]
    lib.cebag_ipc_export.argtypes = [c_void_p, ctypes.c_char_p]
    lib.cebag_ipc_import.argtypes = [ctypes.c_char_p, POINTER(c_void_p)]
    lib.cebag_ipc_close.argtypes = [c_void_p]
    lib.cebag_peer_barrier.argtypes = [POINTER(Exchange), c_int32, c_void_p, c_void_p, c_void_p]
    lib.cebag_prepare_workspace_bytes.argtypes = [POINTER(Table), c_int64]
    lib.cebag_prepare_workspace_bytes.restype = c_size_t
    lib.cebag_prepare_ids.argtypes = [POINTER(Table), c_void_p, c_int64, c_void_p, POINTER(Workspace),
                                      POINTER(PrepareStats), c_void_p]
    lib.cebag_prepare_ids_async.argtypes = [POINTER(Table), c_void_p, c_int64, c_void_p, POINTER(Workspace),
                                            c_void_p, c_void_p]
    lib.cebag_prepare_result_status.argtypes = [POINTER(Table), c_void_p, POINTER(PrepareStats)]
    lib.cebag_available_rows.argtypes = [POINTER(Table), POINTER(c_int64), c_void_p]
    lib.cebag_flush.argtypes = [POINTER(Table), POINTER(Workspace), POINTER(c_int64), c_void_p]
    lib.cebag_preload.argtypes = [POINTER(Table), c_void_p, c_void_p, c_int64, c_void_p]
    lib.cebag_admit_row.argtypes = [POINTER(Table), c_int64, c_int64, c_void_p]
    lib.cebag_evict_slot.argtypes = [POINTER(Table), c_int64, c_void_p]
    lib.cebag_bag_forward.argtypes = [POINTER(BagArgs), c_void_p, c_void_p]
    lib.cebag_backward_workspace_bytes.argtypes = [POINTER(BagArgs)]
    lib.cebag_backward_workspace_bytes.restype = c_size_t
    lib.cebag_bag_backward_fused.argtypes = [POINTER(BagArgs), c_void_p, c_void_p, c_void_p, c_int32, c_float,
                                             c_float, c_void_p, c_size_t, c_int32, c_void_p]
    lib.cebag_bag_backward_plan.argtypes = [POINTER(BagArgs), c_void_p, c_size_t, c_void_p]
    lib.cebag_backward_window_plan_bytes.argtypes = [c_int64]
    lib.cebag_backward_window_plan_bytes.restype = c_size_t
    lib.cebag_bag_backward_plan_window.argtypes = [POINTER(BagArgs), c_int32, c_void_p, c_size_t, POINTER(c_void_p),
                                                   POINTER(c_void_p), POINTER(ctypes.c_uint32), c_void_p]
    lib.cebag_bag_backward_coo.argtypes = [POINTER(BagArgs), c_void_p, c_void_p, c_void_p]
    lib.cebag_bag_backward_dense.argtypes = [POINTER(BagArgs), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.cebag_bag_backward_weights.argtypes = [POINTER(BagArgs), c_void_p, c_void_p, c_void_p]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is c_int and name not in ("cebag_abi_version",):
            fn.restype = c_int


def load():
    """Return the loaded library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CebagError(
            f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            f"(or `python -c 'import __graft_entry__ as g; g.build()'`).  There is no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    _declare(lib)
    if lib.cebag_abi_version() != ABI_VERSION:
        raise CebagError(f"libcebag_b200.so ABI {lib.cebag_abi_version()} != binding ABI {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def launch_count() -> int:
    """Kernels launched by the library since it was loaded."""
    return int(load().cebag_launch_count())


def profile_enable(on: bool):
    check(load().cebag_profile_enable(1 if on else 0))


def profile_collect():
    """{kernel name: (total ms, launch groups)} since the last collect; waits for the recorded events."""
    lib = load()
    k = lib.cebag_profile_num_kernels()
    ms = (ctypes.c_double * k)()
    cnt = (c_int64 * k)()
    check(lib.cebag_profile_collect(ms, cnt))
    return {lib.cebag_profile_kernel_name(i).decode(): (ms[i], int(cnt[i])) for i in range(k) if cnt[i]}


def last_error() -> str:
    return load().cebag_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_CAPACITY:
        raise CacheCapacityError(msg)
    if rc == ERR_INDEX:
        raise IndexError(msg)
    if rc == ERR_INVALID:
        raise ValueError(msg)
    raise CebagError(msg)


__all__ = ["load", "check", "last_error", "Table", "Workspace", "PrepareStats", "PrepareResult", "BagArgs", "byref", "CebagError",
           "CacheCapacityError", "EXPORTS", "LIB_PATH"]
