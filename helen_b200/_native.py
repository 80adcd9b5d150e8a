"""ctypes binding of include/helen_b200.h (the drop-in boundary).

Fails loudly: if the library has not been built, or a call returns an error status, a
RuntimeError / ValueError carrying hb_last_error() is raised.  No fallback path exists.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint8, c_void_p

# HB_LIB: another build of the same library (A/B measurements of compile-time variants, tools/build_variants.py)
LIB_PATH = os.environ.get("HB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libhelen_b200.so")

HB_ABI_VERSION = 4
HB_OK = 0
HB_ERR_INVALID_ARGUMENT = -1
HB_ERR_UNSUPPORTED_DEVICE = -2
HB_ERR_CUDA = -3
HB_ERR_WORKSPACE = -4
HB_ERR_OUT_OF_MEMORY = -5

ENGINE_DEFAULT, ENGINE_FP32, ENGINE_TENSOR = 0, 1, 2
ENGINES = {"default": ENGINE_DEFAULT, "fp32": ENGINE_FP32, "tensor": ENGINE_TENSOR}

_FP = POINTER(c_float)
_U8 = POINTER(c_uint8)


class hb_gru_weights(ctypes.Structure):
    _fields_ = [("weight_ih", _FP * 2), ("weight_hh", _FP * 2), ("bias_ih", _FP * 2), ("bias_hh", _FP * 2)]


class hb_weights(ctypes.Structure):
    _fields_ = [("encoder", hb_gru_weights), ("decoder", hb_gru_weights),
                ("base_weight", _FP), ("base_bias", _FP), ("rle_weight", _FP), ("rle_bias", _FP)]


class hb_launch_plan(ctypes.Structure):
    _fields_ = [("chunkloop", c_int), ("windows_per_cta", c_int), ("stacked_operand", c_int),
                ("recurrence_ctas", c_int), ("projection_workers", c_int), ("heads_workers", c_int),
                ("cooperative", c_int), ("launches", c_int)]


# every symbol include/helen_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hb_abi_version": (c_int, []),
    "hb_last_error": (c_char_p, []),
    "hb_device_count": (c_int, []),
    "hb_create": (c_int, [POINTER(hb_weights), c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "hb_destroy": (None, [c_void_p]),
    "hb_set_engine": (c_int, [c_void_p, c_int]),
    "hb_get_engine": (c_int, [c_void_p]),
    "hb_workspace_bytes": (c_int, [c_void_p, c_int64, c_int, c_int, POINTER(c_size_t)]),
    "hb_predict_windows": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hb_predict_windows_host": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "hb_forward_chunk": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "hb_train_workspace_bytes": (c_int, [c_void_p, c_int64, c_int, POINTER(c_size_t)]),
    "hb_train_step_chunk": (c_int, [c_void_p, POINTER(hb_weights), POINTER(hb_weights), c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hb_launch_count": (c_int64, [c_void_p]),
    "hb_last_launch_plan": (c_int, [c_void_p, POINTER(hb_launch_plan)]),
    "hb_enable_kernel_timing": (c_int, [c_void_p, c_int]),
    "hb_kernel_time_ms": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64), c_int]),
    "hb_dominant_kernel_time_ms": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64), c_int]),
}

_lib = None


def load():
    """Load libhelen_b200.so and bind every declared symbol; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} has not been built (run `python -m helen_b200.build` or __graft_entry__.build()); "
            "helen_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.hb_abi_version() != HB_ABI_VERSION:
        raise RuntimeError(f"libhelen_b200.so ABI {lib.hb_abi_version()} != binding {HB_ABI_VERSION}")
    _lib = lib
    return lib


def last_error():
    return load().hb_last_error().decode("utf-8", "replace")


def check(status):
    """Translate an hb_status into the reference's error behaviour (Python exceptions)."""
    if status >= 0:
        return status
    msg = last_error()
    if status == HB_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if status == HB_ERR_OUT_OF_MEMORY:
        raise MemoryError(msg)
    raise RuntimeError(msg)
