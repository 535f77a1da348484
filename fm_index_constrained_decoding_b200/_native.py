"""ctypes binding of libfmidx.so (include/fmidx.h).  No fallback: a missing library raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfmidx.so")

u64, i64, u32, i32, p, sz = C.c_uint64, C.c_int64, C.c_uint32, C.c_int32, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); every symbol include/fmidx.h declares
SIGNATURES = {
    "fmidx_last_error": (C.c_char_p, []),
    "fmidx_abi_version": (C.c_int, []),
    "fmidx_new": (C.c_int, [C.c_int, C.POINTER(p)]),
    "fmidx_free": (None, [p]),
    "fmidx_set_options": (C.c_int, [p, u32, u32, C.c_int]),
    "fmidx_tune": (C.c_int, [p, C.c_char_p, i64]),
    "fmidx_initialize": (C.c_int, [p, p, sz]),
    "fmidx_initialize_from_file": (C.c_int, [p, C.c_char_p, C.c_int]),
    "fmidx_initialize_device": (C.c_int, [p, p, sz, p]),
    "fmidx_save": (C.c_int, [p, C.c_char_p]),
    "fmidx_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(p)]),
    "fmidx_size": (C.c_int, [p, C.POINTER(u64)]),
    "fmidx_backward_search_step": (C.c_int, [p, u64, u64, u64, C.POINTER(u64)]),
    "fmidx_backward_search_multi": (C.c_int, [p, p, sz, C.POINTER(u64)]),
    "fmidx_distinct": (C.c_int, [p, u64, u64, p, sz, C.POINTER(sz)]),
    "fmidx_distinct_count": (C.c_int, [p, u64, u64, p, sz, C.POINTER(sz)]),
    "fmidx_distinct_count_multi": (C.c_int, [p, p, p, sz, p, sz, p]),
    "fmidx_locate": (C.c_int, [p, u64, C.POINTER(u64)]),
    "fmidx_extract_text": (C.c_int, [p, u64, u64, p, sz, C.POINTER(sz)]),
    "fmidx_search_batch": (C.c_int, [p, p, p, sz, i64, p, p, p]),
    "fmidx_step_batch": (C.c_int, [p, p, p, p, sz, p, p, p]),
    "fmidx_decode_step": (C.c_int, [p, p, i64, p, p, p, p, p, sz, i64, i64, i64, i64, p, p, p, p]),
    "fmidx_distinct_mask_batch": (C.c_int, [p, p, p, sz, i64, u32, p, p, p]),
    "fmidx_query_batch": (C.c_int, [p, C.c_int, p, p, sz, i64, u32, p, p, p, p]),
    "fmidx_distinct_mask_batch_occ": (C.c_int, [p, p, p, sz, i64, u32, p, p, p, p]),
    "fmidx_shard_create": (C.c_int, [p, u32, u32, sz, u32, u32, C.POINTER(p)]),
    "fmidx_shard_export": (C.c_int, [p, u32, C.c_int, p]),
    "fmidx_shard_import": (C.c_int, [p, u32, C.c_int, u32, p]),
    "fmidx_shard_ready": (C.c_int, [p]),
    "fmidx_shard_step": (C.c_int, [p, u32, p, p, i64, p, p, p]),
    "fmidx_shard_graph_prepare": (C.c_int, [p, u32, p, p, i64, p, p]),
    "fmidx_shard_graph_launch": (C.c_int, [p, u32]),
    "fmidx_shard_result": (C.c_int, [p, u32, C.c_int, C.POINTER(p), C.POINTER(p), C.POINTER(p)]),
    "fmidx_shard_submit": (C.c_int, [p, u32, p, p, i64, p, p, p, p]),
    "fmidx_shard_wait": (C.c_int, [p, u32]),
    "fmidx_shard_destroy": (None, [p]),
    "fmidx_ipc_alloc": (C.c_int, [p, sz, C.POINTER(p), p]),
    "fmidx_ipc_open": (C.c_int, [p, p, C.POINTER(p)]),
    "fmidx_ipc_close": (C.c_int, [p, p]),
    "fmidx_ipc_free": (C.c_int, [p, p]),
    "fmidx_mask_logits": (C.c_int, [p, p, p, sz, i64, u32, p, p, C.c_int, i64, p, i32, p, u32, p, p]),
    "fmidx_boost_logits": (C.c_int, [p, p, p, C.c_int, i64, sz, u32, p, p, p, p, p, p, p, p, i32, p]),
    "fmidx_mask_or": (C.c_int, [p, p, u32, sz, p, p]),
    "fmidx_mask_popcount": (C.c_int, [p, p, sz, u32, p, p]),
    "fmidx_mask_compact": (C.c_int, [p, p, p, sz, u32, p, p, p, p]),
    "fmidx_locate_batch": (C.c_int, [p, p, sz, p, p]),
    "fmidx_extract_batch": (C.c_int, [p, p, p, sz, u64, p, p, p]),
    "fmidx_search_host": (C.c_int, [p, p, p, sz, i64, p, p]),
    "fmidx_query_host": (C.c_int, [p, p, p, sz, i64, u32, p, p, p]),
    "fmidx_query_host_summary": (C.c_int, [p, p, p, sz, i64, u32, p, p, p]),
    "fmidx_query_submit": (C.c_int, [p, C.c_int, p, p, sz, i64, u32, p, p, p]),
    "fmidx_query_wait": (C.c_int, [p, C.c_int]),
    "fmidx_query_graph_prepare": (C.c_int, [p, C.c_int, sz, sz, i64, u32, C.POINTER(p), C.POINTER(p)]),
    "fmidx_query_graph_launch": (C.c_int, [p, C.c_int]),
    "fmidx_query_graph_prepare_device": (C.c_int, [p, C.c_int, p, p, sz, i64, u32, p, p, p]),
    "fmidx_query_graph_launch_on": (C.c_int, [p, C.c_int, p]),
    "fmidx_locate_host": (C.c_int, [p, p, sz, p]),
    "fmidx_info": (C.c_int, [p, C.POINTER(u64)]),
    "fmidx_debug_copy_sa": (C.c_int, [p, p, sz]),
    "fmidx_gather_peak": (C.c_int, [p, C.c_int, sz, C.POINTER(C.c_double)]),
    "fmidx_launch_count": (u64, []),
}

FMIDX_EZERO = -6
FMIDX_ERANGE = -5
DTYPE_CODES = {"float32": 0, "bfloat16": 1, "float16": 2}

_LIB = None


class FmidxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fmidx error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Loads libfmidx.so; raises ImportError (never falls back) when it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libfmidx.so is not built (%s). Run `python -m fm_index_constrained_decoding_b200.build` "
                "(needs nvcc). This package has no CPU or PyTorch fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise FmidxError(rc, lib().fmidx_last_error().decode("utf-8", "replace"))


def default_device():
    return int(os.environ.get("FMIDX_DEVICE", os.environ.get("LOCAL_RANK", "0")))
