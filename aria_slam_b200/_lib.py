"""ctypes binding of libaslam_b200.so (the C ABI of include/aslam_b200.h).

There is no CPU fallback: if the shared library is missing the import of any product class
raises, and every entry point fails with ASLAM_E_NODEVICE when no CUDA device is visible.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ASLAM_B200_LIB") or os.path.join(_HERE, "libaslam_b200.so")     # env override: A/B builds

OK, E_INVALID, E_CUDA, E_NOMEM, E_STATE, E_OVERFLOW, E_NODEVICE = range(7)

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])
MATCH_DTYPE = np.dtype([("query_idx", "<i4"), ("train_idx", "<i4"), ("distance", "<f4")])
BOX_DTYPE = np.dtype([("x1", "<f4"), ("y1", "<f4"), ("x2", "<f4"), ("y2", "<f4"),
                      ("confidence", "<f4"), ("class_id", "<i4")])


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("edge_threshold", C.c_int32), ("first_level", C.c_int32), ("wta_k", C.c_int32),
                ("score_type", C.c_int32), ("patch_size", C.c_int32), ("fast_threshold", C.c_int32),
                ("blur_mode", C.c_int32), ("grid_cell", C.c_int32)]


class AslamError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("aslam_b200 error %d: %s" % (code, msg))
        self.code = code


# name -> (restype, argtypes); every symbol declared in include/aslam_b200.h
_P = C.c_void_p
SYMBOLS = {
    "aslam_device_count": (C.c_int, []),
    "aslam_debug_bounds": (C.c_int, [C.POINTER(C.c_uint64)]),
    "aslam_ctx_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "aslam_ctx_destroy": (None, [_P]),
    "aslam_ctx_sync": (C.c_int, [_P]),
    "aslam_ctx_stream": (_P, [_P]),
    "aslam_ctx_sm_count": (C.c_int, [_P]),
    "aslam_last_error": (C.c_char_p, []),
    "aslam_version": (C.c_char_p, []),
    "aslam_kernel_launches": (C.c_uint64, []),
    "aslam_prof_enable": (C.c_int, [_P, C.c_int]),
    "aslam_prof_count": (C.c_int, []),
    "aslam_prof_name": (C.c_char_p, [C.c_int]),
    "aslam_prof_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "aslam_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "aslam_host_alloc_wc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "aslam_host_free": (C.c_int, [_P]),
    "aslam_memcpy_async": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "aslam_orb_default_params": (None, [C.POINTER(OrbParams)]),
    "aslam_orb_create": (C.c_int, [_P, C.POINTER(OrbParams), C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "aslam_orb_destroy": (None, [_P]),
    "aslam_orb_set_max_features": (C.c_int, [_P, C.c_int]),
    "aslam_orb_get_max_features": (C.c_int, [_P]),
    "aslam_orb_capacity": (C.c_int, [_P]),
    "aslam_orb_extract_async": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int]),
    "aslam_orb_extract_batch_async": (C.c_int, [_P, _P, C.c_int, C.c_size_t, C.c_int, _P]),
    "aslam_orb_sync": (C.c_int, [_P, _P, _P, _P]),
    "aslam_orb_extract_batch_device": (C.c_int, [_P, _P, C.c_int, C.c_size_t, C.c_int, _P]),
    "aslam_orb_device_results": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "aslam_orb_track_batch_async": (C.c_int, [_P, _P, C.c_int, C.c_size_t, C.c_int, C.c_float, C.c_int]),
    "aslam_orb_track_sync": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "aslam_orb_level_size": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "aslam_orb_read_level": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "aslam_orb_read_corners": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_int)]),
    "aslam_match_knn2": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_float, _P, C.POINTER(C.c_int)]),
    "aslam_match_knn2_dynamic": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, C.c_int, C.c_float, _P, C.c_int, C.c_int, _P,
                                           C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "aslam_knn2_raw": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P, _P]),
    "aslam_match_multi": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, C.c_float, _P, _P]),
    "aslam_match_knn2_resident": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_float, _P, C.POINTER(C.c_int)]),
    "aslam_orb_retain_descriptors": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int32)]),
    "aslam_match_knn2_device": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_float, _P, _P]),
    "aslam_loopdb_create": (C.c_int, [_P, C.c_int64, C.c_int, C.POINTER(_P)]),
    "aslam_loopdb_destroy": (None, [_P]),
    "aslam_loopdb_add": (C.c_int, [_P, _P, C.c_int, C.c_uint64]),
    "aslam_loopdb_add_at": (C.c_int, [_P, _P, C.c_int, C.c_uint64, C.c_int64]),
    "aslam_loopdb_add_device": (C.c_int, [_P, _P, C.c_int, C.c_uint64]),
    "aslam_loopdb_clear": (C.c_int, [_P]),
    "aslam_loopdb_save": (C.c_int, [_P, C.c_char_p]),
    "aslam_loopdb_load": (C.c_int, [_P, C.c_char_p]),
    "aslam_loopdb_num_keyframes": (C.c_int, [_P]),
    "aslam_loopdb_num_rows": (C.c_int64, [_P]),
    "aslam_loopdb_set_row_base": (C.c_int, [_P, C.c_int64]),
    "aslam_loopdb_keyframes": (C.c_int, [_P, _P, _P]),
    "aslam_loopdb_query_segmented": (C.c_int, [_P, _P, C.c_int, C.c_double, _P]),
    "aslam_loopdb_query_global_device": (C.c_int, [_P, _P, C.c_int, _P]),
    "aslam_loopdb_query_global_push_device": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int]),
    "aslam_loopdb_query_global_exchange_device": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_uint64, _P]),
    "aslam_loopdb_exchange_status": (C.c_int, [_P, C.POINTER(C.c_uint32), _P, C.c_int, C.c_int]),
    "aslam_loopdb_set_pipelining": (C.c_int, [_P, C.c_int]),
    "aslam_loopdb_capacity_rows": (C.c_int64, [_P]),
    "aslam_loopdb_query_segmented_device": (C.c_int, [_P, _P, C.c_int, C.c_double, _P]),
    "aslam_loopdb_match_keyframe": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_double, _P, C.POINTER(C.c_int)]),
    "aslam_loopdb_shard_handle_bytes": (C.c_size_t, []),
    "aslam_loopdb_shard_export": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "aslam_loopdb_shard_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "aslam_loopdb_query_global_sharded_device": (C.c_int, [_P, _P, C.c_int, _P]),
    "aslam_loopdb_shard_group_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_int64, C.c_int, C.POINTER(_P)]),
    "aslam_loopdb_shard_group_destroy": (None, [_P]),
    "aslam_loopdb_shard_group_num_shards": (C.c_int, [_P]),
    "aslam_loopdb_shard_group_add": (C.c_int, [_P, _P, C.c_int, C.c_uint64]),
    "aslam_loopdb_shard_group_clear": (C.c_int, [_P]),
    "aslam_loopdb_shard_group_num_keyframes": (C.c_int, [_P]),
    "aslam_loopdb_shard_group_num_rows": (C.c_int64, [_P]),
    "aslam_loopdb_shard_group_keyframes": (C.c_int, [_P, _P, _P, _P]),
    "aslam_loopdb_shard_group_query_segmented": (C.c_int, [_P, _P, C.c_int, C.c_double, _P]),
    "aslam_loopdb_shard_group_query_global": (C.c_int, [_P, _P, C.c_int, _P]),
    "aslam_loopdb_shard_group_read_result": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "aslam_loopdb_shard_group_match_keyframe": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_double, _P, C.POINTER(C.c_int)]),
    "aslam_loop_exchange_buffer_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "aslam_loopdb_query_global": (C.c_int, [_P, _P, C.c_int, _P]),
    "aslam_loop_merge_top2_device": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "aslam_microbench": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built: the product path must
    fail loudly rather than fall back to anything else."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libaslam_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                              "this path has no fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, allow_overflow=False):
    if rc == OK or (allow_overflow and rc == E_OVERFLOW):
        return rc
    raise AslamError(rc, lib().aslam_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Pointer of a numpy array, a torch tensor (host or device) or None / int."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError("cannot take a pointer of %r" % type(a))


class Context:
    """aslam_ctx: one CUDA device + one stream (borrowed or private)."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        check(lib().aslam_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device

    def sync(self):
        check(lib().aslam_ctx_sync(self.h))

    @property
    def stream(self):
        return lib().aslam_ctx_stream(self.h)

    @property
    def sm_count(self):
        return lib().aslam_ctx_sm_count(self.h)

    def microbench(self, which):
        v = C.c_double()
        check(lib().aslam_microbench(self.h, which, C.byref(v)))
        return v.value

    def prof_enable(self, on=True):
        check(lib().aslam_prof_enable(self.h, 1 if on else 0))

    def prof_read(self, reset=True):
        """{kernel kind: (total device ms, launches)} accumulated since the last reset."""
        out = {}
        for k in range(lib().aslam_prof_count()):
            ms, n = C.c_double(), C.c_int64()
            check(lib().aslam_prof_read(self.h, k, C.byref(ms), C.byref(n), 1 if reset else 0))
            out[lib().aslam_prof_name(k).decode()] = (ms.value, n.value)
        return out

    def close(self):
        if self.h:
            lib().aslam_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def kernel_launches():
    return int(lib().aslam_kernel_launches())
