"""TEST INFRASTRUCTURE ONLY.  The reference's own Smith-Waterman (helen/modules/src/local_reassembly/ssw.c,
ssw_cpp.cpp), compiled from /root/reference by build_ref.py into oracle/_ref/libssw_ref.so, behind a ctypes
call.  It is the checker for helen_b200's stitch library (tests/test_stitch.py) and the CPU baseline of
tests/bench_stitch.py; nothing under helen_b200/ may import it."""
import ctypes

from . import build_ref


class _Result(ctypes.Structure):
    _fields_ = [(name, ctypes.c_int32) for name in ("score", "score2", "ref_begin", "ref_end", "query_begin",
                                                    "query_end", "ref_end2", "mismatches", "cigar_len")]


_lib = None


def load():
    """None when neither the prebuilt library nor the reference sources are available."""
    global _lib
    if _lib is None:
        path = build_ref.build(pybind=False)
        if path is None:
            return None
        _lib = ctypes.CDLL(path)
        _lib.ref_ssw_align.restype = ctypes.c_int
        _lib.ref_ssw_align.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_Result), ctypes.c_char_p, ctypes.c_int32]
    return _lib


def align(ref, query, match=4, mismatch=6, gap_open=8, gap_extend=2):
    """Stitch.py:110-135's call -> dict(score, ref_begin, ref_end, query_begin, query_end, mismatches, cigar)."""
    lib = load()
    out = _Result()
    cigar = ctypes.create_string_buffer(16 * (len(ref) + len(query)) + 64)
    rc = lib.ref_ssw_align(ref.encode(), len(ref), query.encode(), match, mismatch, gap_open, gap_extend,
                           ctypes.byref(out), cigar, len(cigar))
    if rc == 1 or (rc == 0 and out.score == 0):
        return dict(score=0)       # nothing aligned: the other fields are leftovers (Stitch.py:138 reads none of them)
    if rc:
        raise RuntimeError("ref_ssw_align failed: %d" % rc)
    return dict(score=out.score, ref_begin=out.ref_begin, ref_end=out.ref_end, query_begin=out.query_begin,
                query_end=out.query_end, mismatches=out.mismatches, cigar=cigar.value.decode())
