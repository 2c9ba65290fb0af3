"""ctypes loader for libgofeature_b200.so (the C ABI of include/gofeature_b200.h).

There is no fallback: if the shared library is missing this module raises at import of
`lib()`; if no CUDA device is present `gf_context_create` fails.  Nothing here touches
oracle/.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgofeature_b200.so")

c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_u32 = ctypes.c_uint32
c_f32 = ctypes.c_float
c_vp = ctypes.c_void_p
c_cp = ctypes.c_char_p
P = ctypes.POINTER


class gf_stats(ctypes.Structure):
    _fields_ = [(n, c_u64) for n in ("searches", "scan_launches", "gemm_launches", "merge_launches",
                                     "other_launches", "rows_scanned", "quirk_fallbacks", "coalesced_batches",
                                     "h2d_bytes", "d2h_bytes", "scan_kernel_ns", "scan_kernel_timed",
                                     "tensor_passes", "uncertified_queries", "gemm_kernel_ns", "gemm_kernel_timed",
                                     "host_stage_ns", "host_launch_ns", "host_wait_ns", "host_resolve_ns",
                                     "host_calls", "add_stage_ns", "add_place_ns", "add_index_ns", "add_rows",
                                     "tier2_queries")]


# name -> (restype, argtypes); every symbol include/gofeature_b200.h declares
SIGNATURES = {
    "gf_strerror": (c_cp, [c_int]),
    "gf_status_name": (c_cp, [c_int]),
    "gf_last_error_detail": (c_cp, []),
    "gf_version": (c_cp, []),
    "gf_util_normalize_float32": (c_int, [c_vp, c_vp, c_i64]),
    "gf_util_transpose_1d": (c_int, [c_int, c_vp, c_i64, c_i64, c_vp]),
    "gf_device_count": (c_int, [P(c_int)]),
    "gf_device_total_mem": (c_int, [c_int, P(c_u64)]),
    "gf_context_create": (c_int, [c_int, P(c_vp)]),
    "gf_context_create_multi": (c_int, [P(c_int), c_int, P(c_vp)]),
    "gf_context_device_count": (c_int, [c_vp]),
    "gf_context_destroy": (c_int, [c_vp]),
    "gf_context_device": (c_int, [c_vp]),
    "gf_context_synchronize": (c_int, [c_vp]),
    "gf_buffer_create": (c_int, [c_vp, c_u64, P(c_vp)]),
    "gf_buffer_wrap": (c_int, [c_vp, c_vp, c_u64, P(c_vp)]),
    "gf_buffer_write": (c_int, [c_vp, c_vp, c_u64]),
    "gf_buffer_read": (c_int, [c_vp, c_vp, c_u64]),
    "gf_buffer_copy": (c_int, [c_vp, c_vp]),
    "gf_buffer_reset": (c_int, [c_vp]),
    "gf_buffer_slice": (c_int, [c_vp, c_i64, c_i64, P(c_vp)]),
    "gf_buffer_size": (c_u64, [c_vp]),
    "gf_buffer_device_ptr": (c_vp, [c_vp]),
    "gf_buffer_free": (c_int, [c_vp]),
    "gf_cache_create": (c_int, [c_vp, c_int, c_i64, P(c_vp)]),
    "gf_cache_create_ex": (c_int, [c_vp, c_int, c_i64, ctypes.c_uint, P(c_vp)]),
    "gf_cache_has_shadow": (c_int, [c_vp]),
    "gf_cache_destroy": (c_int, [c_vp]),
    "gf_cache_new_set": (c_int, [c_vp, c_cp, c_int, c_int, c_int]),
    "gf_cache_destroy_set": (c_int, [c_vp, c_cp]),
    "gf_cache_get_set": (c_int, [c_vp, c_cp, P(c_vp)]),
    "gf_cache_get_block_size": (c_i64, [c_vp]),
    "gf_cache_get_empty_block": (c_int, [c_vp, c_int, P(c_vp)]),
    "gf_cache_block_count": (c_int, [c_vp]),
    "gf_cache_block_at": (c_int, [c_vp, c_int, P(c_vp)]),
    "gf_block_capacity": (c_int, [c_vp]),
    "gf_block_margin": (c_int, [c_vp]),
    "gf_block_is_owned": (c_int, [c_vp]),
    "gf_block_next_index": (c_int, [c_vp]),
    "gf_block_acquire": (c_int, [c_vp, c_cp, c_int, c_int, c_int]),
    "gf_block_release": (c_int, [c_vp]),
    "gf_block_insert": (c_int, [c_vp, c_i64, c_vp, c_u64, c_vp, c_vp]),
    "gf_block_delete": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, P(c_i64)]),
    "gf_block_search": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, P(c_vp)]),
    "gf_set_add": (c_int, [c_vp, c_i64, c_vp, c_u64, c_vp, c_vp]),
    "gf_set_delete": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, P(c_i64)]),
    "gf_set_update": (c_int, [c_vp, c_i64, c_vp, c_u64, c_vp, c_vp, c_vp, P(c_i64)]),
    "gf_set_read": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_u64, c_vp, P(c_i64)]),
    "gf_set_search": (c_int, [c_vp, c_f32, c_int, c_int, c_vp, c_u64, P(c_vp)]),
    "gf_set_destroy": (c_int, [c_vp]),
    "gf_set_compact": (c_int, [c_vp, P(c_i64)]),
    "gf_set_export": (c_int, [c_vp, c_vp, c_u64, c_vp, c_u64, c_vp, P(c_i64), P(c_u64)]),
    "gf_set_dims": (c_int, [c_vp]),
    "gf_set_precision": (c_int, [c_vp]),
    "gf_set_batch": (c_int, [c_vp]),
    "gf_set_block_count": (c_int, [c_vp]),
    "gf_set_scanned_rows": (c_i64, [c_vp]),
    "gf_set_live_rows": (c_i64, [c_vp]),
    "gf_result_batch": (c_int, [c_vp]),
    "gf_result_count": (c_int, [c_vp, c_int]),
    "gf_result_score": (c_f32, [c_vp, c_int, c_int]),
    "gf_result_id": (c_vp, [c_vp, c_int, c_int, P(c_u32)]),
    "gf_result_slot": (c_u64, [c_vp, c_int, c_int]),
    "gf_result_total": (c_i64, [c_vp]),
    "gf_result_id_bytes": (c_i64, [c_vp]),
    "gf_result_export": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gf_result_free": (c_int, [c_vp]),
    "gf_key_make": (c_u64, [c_f32, c_u32]),
    "gf_key_score": (c_f32, [c_u64]),
    "gf_key_slot": (c_u32, [c_u64]),
    "gf_set_set_shard": (c_int, [c_vp, c_int, c_int]),
    "gf_set_search_device": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    "gf_merge_keys_device": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "gf_set_resolve_keys": (c_int, [c_vp, c_f32, c_int, c_int, c_vp, P(c_vp)]),
    "gf_set_add_synthetic": (c_int, [c_vp, c_i64, c_u64, c_i64, c_int, c_cp]),
    "gf_set_add_synthetic_ex": (c_int, [c_vp, c_i64, c_u64, c_i64, c_int, c_cp, c_u64, c_f32, c_u32]),
    "gf_context_stats": (c_int, [c_vp, P(gf_stats)]),
    "gf_context_reset_stats": (c_int, [c_vp]),
    "gf_context_set_profiling": (c_int, [c_vp, c_int]),
    "gf_set_set_coalescing": (c_int, [c_vp, c_int]),
    "gf_set_set_search_path": (c_int, [c_vp, c_int]),
    "gf_set_search_device_ex": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp]),
    "gf_set_device_counters": (c_int, [c_vp, P(c_u64), P(c_u64)]),
    "gf_exchange_create": (c_int, [c_vp, c_int, c_int, c_int, c_int, P(c_vp)]),
    "gf_exchange_handle": (c_int, [c_vp, c_vp, c_u64]),
    "gf_exchange_connect": (c_int, [c_vp, c_vp, c_u64]),
    "gf_exchange_merge_device": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gf_set_search_device_xchg": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "gf_exchange_trace": (c_int, [c_vp, c_int, c_vp]),
    "gf_exchange_destroy": (c_int, [c_vp]),
}

_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and os.environ.get("GF_NO_AUTOBUILD") is None:
            # a fresh checkout has sources only: compile the CUDA library in-tree (nvcc, sm_100a).  This is a
            # build step, not a fallback -- if it fails the import fails.
            import subprocess
            subprocess.call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], stdout=subprocess.DEVNULL,
                            stderr=subprocess.DEVNULL)
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "gofeature_b200: %s is missing -- build it with __graft_entry__.build() "
                "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the header and the library drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
