"""ctypes view of include/cssm.h and the loader of libcssm_gpu.so.

There is no CPU fallback: if the shared library is missing `lib()` raises, and every compute
entry point of the library itself fails with CSSM_ERR_CUDA when no device is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CSSM_LIB: an alternative build of the SAME library (tuning experiments with other launch bounds); default: the in-tree build
LIB_PATH = os.environ.get("CSSM_LIB") or os.path.join(_HERE, "csrc", "libcssm_gpu.so")

# enums (include/cssm.h)
SDE_BROWNIAN, SDE_GEN_BROWNIAN, SDE_OU = 0, 1, 2
F_FIRST, F_SEASONAL = 0, 1
OBS_POISSON, OBS_NEGBIN, OBS_NORMAL, OBS_BERNOULLI, OBS_LGCP, OBS_STUDENT_T, OBS_ZIP, OBS_BETA = 0, 1, 2, 3, 4, 5, 6, 7
STEP_EXACT, STEP_EULER = 0, 1
RESAMPLE_SYSTEMATIC, RESAMPLE_STRATIFIED, RESAMPLE_MULTINOMIAL = 0, 1, 2
F32, F64 = 0, 1
SERIES_AUTO, SERIES_THREE_LAUNCH, SERIES_SINGLE_LAUNCH = 0, 1, 2
TIE_REFERENCE, TIE_FIRST = 0, 1
SCAN_AUTO, SCAN_EXACT = 0, 1
MAX_RANKS, SHARD_BLOB_BYTES = 8, 1024

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint64_p = C.POINTER(C.c_uint64)
c_uint8_p = C.POINTER(C.c_uint8)


class Leaf(C.Structure):
    _fields_ = [
        ("sde_kind", C.c_int32),
        ("dim", C.c_int32),
        ("f_kind", C.c_int32),
        ("period", C.c_int32),
        ("harmonics", C.c_int32),
        ("m0", c_double_p),
        ("c0", c_double_p),
        ("phi", c_double_p),
        ("mu", c_double_p),
        ("sigma", c_double_p),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_leaves", C.c_int32),
        ("leaves", C.POINTER(Leaf)),
        ("obs_kind", C.c_int32),
        ("has_scale", C.c_int32),
        ("scale", C.c_double),
        ("step_mode", C.c_int32),
        ("lgcp_precision", C.c_int32),
        ("obs_df", C.c_int32),
        ("reserved", C.c_int32),
    ]


class CssmError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the JNI shim rethrows the same way)."""

    def __init__(self, status, message):
        super().__init__(f"cssm status {status}: {message}")
        self.status = status


# name -> (argtypes) ; every function returns int unless listed in _RESTYPES
_FILTER = C.c_void_p
_SIGNATURES = {
    "cssm_version": [],
    "cssm_last_error": [],
    "cssm_device_count": [c_int32_p],
    "cssm_filter_create": [C.POINTER(ModelDesc), C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64,
                           C.c_uint64, C.POINTER(_FILTER)],
    "cssm_filter_create_sharded": [C.POINTER(ModelDesc), C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                   C.c_uint64, C.c_int, C.c_int, C.POINTER(_FILTER)],
    "cssm_filter_shard_export": [_FILTER, C.c_void_p],
    "cssm_filter_shard_connect": [_FILTER, C.c_void_p, C.c_int],
    "cssm_filter_shard_info": [_FILTER, c_int32_p, c_int32_p, c_int64_p],
    "cssm_group_init": [C.POINTER(_FILTER), C.c_int, C.c_double],
    "cssm_group_init_injected": [C.POINTER(_FILTER), C.c_int, C.c_double, c_double_p],
    "cssm_group_step_injected": [C.POINTER(_FILTER), C.c_int, C.c_double, C.c_int, C.c_double, c_double_p, c_double_p,
                                 c_double_p, c_double_p, c_double_p, c_int32_p, c_double_p, c_int32_p],
    "cssm_group_get_particles": [C.POINTER(_FILTER), C.c_int, c_double_p],
    "cssm_group_ll": [C.POINTER(_FILTER), C.c_int, c_double_p, c_double_p, c_uint8_p, C.c_int64, c_double_p,
                      C.POINTER(C.c_float)],
    "cssm_filter_set_params": [_FILTER, C.POINTER(ModelDesc)],
    "cssm_filter_reseed": [_FILTER, C.c_uint64, C.c_uint64],
    "cssm_filter_set_stream": [_FILTER, C.c_void_p],
    "cssm_filter_destroy": [_FILTER],
    "cssm_filter_dim": [_FILTER, c_int32_p],
    "cssm_filter_n_particles": [_FILTER, c_int64_p],
    "cssm_filter_init": [_FILTER, C.c_double],
    "cssm_filter_init_state": [_FILTER, C.c_double, c_double_p],
    "cssm_filter_step": [_FILTER, C.c_double, C.c_int, C.c_double, c_double_p, c_int32_p],
    "cssm_filter_ll": [_FILTER, c_double_p, c_double_p, c_uint8_p, C.c_int64, c_double_p],
    "cssm_filter_load_series": [_FILTER, c_double_p, c_double_p, c_uint8_p, C.c_int64],
    "cssm_filter_ll_resident": [_FILTER, c_double_p, c_double_p, c_int32_p],
    "cssm_filter_series_len": [_FILTER, c_int64_p],
    "cssm_filter_run": [_FILTER, c_double_p, c_double_p, c_uint8_p, C.c_int64, c_double_p, c_double_p],
    "cssm_filter_last_elapsed_ms": [_FILTER, C.POINTER(C.c_float)],
    "cssm_filter_last_launches": [_FILTER, c_int64_p],
    "cssm_filter_series_mode": [_FILTER, C.c_int],
    "cssm_filter_profile": [_FILTER, C.c_int],
    "cssm_filter_profile_read": [_FILTER, c_double_p, c_int64_p],
    "cssm_filter_get_particles": [_FILTER, c_double_p],
    "cssm_filter_sample_one": [_FILTER, c_double_p],
    "cssm_filter_get_ll": [_FILTER, c_double_p, c_int32_p],
    "cssm_filter_mean_state": [_FILTER, c_double_p],
    "cssm_filter_intervals": [_FILTER, C.c_double, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p],
    "cssm_filter_forecast": [_FILTER, C.c_double, C.c_double, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p,
                             c_double_p],
    "cssm_filter_forecast_cloud": [_FILTER, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p],
    "cssm_filter_set_tie_rule": [_FILTER, C.c_int],
    "cssm_filter_scan_mode": [_FILTER, C.c_int],
    "cssm_filter_scan_stats": [_FILTER, c_int64_p, c_int64_p],
    "cssm_filter_paths_enable": [_FILTER, C.c_int64],
    "cssm_filter_paths_len": [_FILTER, c_int64_p],
    "cssm_filter_get_paths": [_FILTER, c_int32_p, C.c_int64, c_double_p],
    "cssm_resample": [C.c_int, c_double_p, C.c_int64, c_double_p, C.c_int64, c_int32_p, C.c_int],
    "cssm_filter_init_injected": [_FILTER, C.c_double, c_double_p],
    "cssm_filter_step_injected": [_FILTER, C.c_double, C.c_int, C.c_double, c_double_p, c_double_p,
                                  c_double_p, c_double_p, c_double_p, c_int32_p, c_double_p, c_int32_p],
    "cssm_filter_n_substeps": [_FILTER, C.c_double, c_int64_p],
}
_RESTYPES = {"cssm_last_error": C.c_char_p}

_lib = None


def declared_symbols():
    """Names this binding expects; tests compare them with include/cssm.h and the .so exports."""
    return sorted(_SIGNATURES)


def lib():
    """Load libcssm_gpu.so (built in-tree by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
                " There is no CPU fallback.")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, args in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = l
    return _lib


def check(status):
    if status != 0:
        msg = lib().cssm_last_error()
        raise CssmError(status, msg.decode("utf-8", "replace") if msg else "")


def dptr(a):
    """numpy float64 C-contiguous array -> double*; None -> NULL."""
    if a is None:
        return None
    return a.ctypes.data_as(c_double_p)
