"""ctypes binding of include/helen_h5write.h (the prediction-file writer inside libhelen_feed.so)."""
import ctypes
import os
from ctypes import POINTER, c_char, c_char_p, c_int, c_int64, c_uint64, c_void_p

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libhelen_feed.so")

HW_ABI_VERSION = 1
HW_OK, HW_E_EXISTS, HW_E_TYPE, HW_E_IO, HW_E_ARGUMENT = 0, 1, 2, 3, 4

# every symbol include/helen_h5write.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hw_abi_version": (c_int, []),
    "hw_create": (c_int, [c_char_p, POINTER(c_void_p), c_char_p, c_int]),
    "hw_dataset": (c_int, [c_void_p, c_char_p, c_char, c_int, c_int, POINTER(c_uint64), c_void_p, c_char_p, c_int]),
    "hw_rows": (c_int, [c_void_p, c_char_p, c_int64, c_char_p, c_char, c_int, c_int, POINTER(c_uint64), c_void_p, c_char_p, c_int]),
    "hw_contains": (c_int, [c_void_p, c_char_p]),
    "hw_close": (c_int, [c_void_p, c_char_p, c_int]),
}

_lib = None


def load():
    """Load libhelen_feed.so and bind the writer's symbols; raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} has not been built (run `python -m helen_b200.build` or __graft_entry__.build())")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = restype, argtypes
    if lib.hw_abi_version() != HW_ABI_VERSION:
        raise RuntimeError(f"{LIB_PATH}: ABI version {lib.hw_abi_version()}, binding expects {HW_ABI_VERSION}")
    _lib = lib
    return lib


def raise_for(status, message):
    if status == HW_E_EXISTS:
        raise ValueError(message)
    if status == HW_E_TYPE:
        raise TypeError("minih5 " + message)
    if status == HW_E_ARGUMENT:
        raise ValueError(message)
    raise IOError(message)
