"""ctypes binding of include/helen_stitch.h (host-side stitch library).

Fails loudly like _native.py: a missing library or an error status raises; there is no Python
implementation of the alignment behind it.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int32, c_int64, c_uint8, c_void_p

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libhelen_stitch.so")

HS_ABI_VERSION = 1
HS_E_ARGUMENT, HS_E_CAPACITY, HS_E_RANGE, HS_E_TRACE, HS_E_CIGAR = -1, -2, -3, -4, -5
HS_WARN_NO_ALIGNMENT, HS_WARN_NO_ANCHOR, HS_WARN_NO_OVERLAP = 0, 1, 2


class hs_scoring(ctypes.Structure):
    _fields_ = [("match", c_int32), ("mismatch", c_int32), ("gap_open", c_int32), ("gap_extend", c_int32)]


class hs_alignment(ctypes.Structure):
    _fields_ = [("score", c_int32), ("ref_begin", c_int32), ("ref_end", c_int32), ("query_begin", c_int32),
                ("query_end", c_int32), ("mismatches", c_int32), ("cigar_len", c_int32), ("kernel", c_int32)]


# every symbol include/helen_stitch.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hs_abi_version": (c_int32, []),
    "hs_last_error": (c_char_p, []),
    "hs_ssw_align": (c_int32, [c_char_p, c_int32, c_char_p, c_int32, POINTER(hs_scoring), POINTER(hs_alignment),
                               c_char_p, c_int32]),
    "hs_anchor_from_cigar": (c_int32, [c_char_p, c_int32, c_int32, POINTER(c_int32), POINTER(c_int32)]),
    "hs_decode_region": (c_int64, [POINTER(c_int64), POINTER(c_uint8), POINTER(c_uint8), c_int64, c_char_p, c_int64]),
    "hs_stitcher_create": (c_void_p, [POINTER(hs_scoring), c_int32, c_double]),
    "hs_stitcher_destroy": (None, [c_void_p]),
    "hs_stitcher_add": (c_int32, [c_void_p, c_int64, c_int64, c_char_p, c_int64]),
    "hs_stitcher_run": (c_int32, [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64),
                                  POINTER(c_int64)]),
    "hs_stitcher_sequence": (c_int64, [c_void_p, c_char_p, c_int64]),
}

_lib = None


def load():
    """Load libhelen_stitch.so and bind every declared symbol; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} has not been built (run `python -m helen_b200.build` or __graft_entry__.build())")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = restype, argtypes
    if lib.hs_abi_version() != HS_ABI_VERSION:
        raise RuntimeError(f"{LIB_PATH}: ABI version {lib.hs_abi_version()}, binding expects {HS_ABI_VERSION}")
    _lib = lib
    return lib


def check(status):
    """Raise on a negative status, carrying hs_last_error(); returns the status otherwise."""
    if status is not None and status < 0:
        message = load().hs_last_error().decode()
        raise (ValueError if status in (HS_E_ARGUMENT, HS_E_CIGAR) else RuntimeError)(f"helen_stitch error {status}: {message}")
    return status
