"""Stitch with the reference's interface (helen/modules/python/Stitch.py:14-301), computed by the host
library behind include/helen_stitch.h.

``Aligner`` / ``Alignment`` / ``Filter`` stand where the reference's pybind module ``helen.build.HELEN``
stands (pybind_api.h:16-47); ``Stitch`` keeps the reference's method names and argument meaning.
"""
import concurrent.futures
import ctypes
import sys

import numpy as np

from . import _stitch_native as native
from . import _feed_native
from .DataStore import open_predictions, region_reader
from .FileManager import FileManager
from .options import StitchOptions
from .TextColor import TextColor


class Alignment(object):
    """Fields of HELEN.Alignment that Align_cpp fills (pybind_api.h:18-30)."""

    def __init__(self):
        self.Clear()

    def Clear(self):
        self.best_score = 0
        self.best_score2 = 0
        self.reference_begin = 0
        self.reference_end = 0
        self.query_begin = 0
        self.query_end = 0
        self.ref_end_next_best = 0
        self.mismatches = 0
        self.cigar_string = ""


class Filter(object):
    """HELEN.Filter (pybind_api.h:33-39).  Only the default filter (begin position and cigar always reported,
    which is what Stitch.py:112 constructs) is supported."""

    def __init__(self, report_begin_position=True, report_cigar=True, score_filter=0, distance_filter=32767):
        self.report_begin_position = report_begin_position
        self.report_cigar = report_cigar
        self.score_filter = score_filter
        self.distance_filter = distance_filter


class Aligner(object):
    """HELEN.Aligner (pybind_api.h:42-46): SetReferenceSequence + Align_cpp."""

    def __init__(self, match_score=2, mismatch_penalty=2, gap_opening_penalty=3, gap_extending_penalty=1):
        self._scoring = native.hs_scoring(match_score, mismatch_penalty, gap_opening_penalty, gap_extending_penalty)
        self._reference = b""
        self._lib = native.load()

    def SetReferenceSequence(self, seq, length):
        self._reference = seq.encode()[:length]
        return len(self._reference)

    def Align_cpp(self, query, filter, alignment, maskLen):
        if not (filter.report_begin_position and filter.report_cigar and filter.score_filter == 0
                and filter.distance_filter == 32767) or maskLen >= 15:
            raise ValueError("helen_b200 Aligner supports the default Filter and maskLen < 15 only (Stitch.py:112,135)")
        alignment.Clear()
        query = query.encode()
        if len(self._reference) == 0 or len(query) == 0:
            return False                                    # ssw_cpp.cpp:324-327
        out = native.hs_alignment()
        cigar = ctypes.create_string_buffer(16 * (len(self._reference) + len(query)) + 64)
        native.check(self._lib.hs_ssw_align(self._reference, len(self._reference), query, len(query),
                                            ctypes.byref(self._scoring), ctypes.byref(out), cigar, len(cigar)))
        alignment.best_score = out.score
        alignment.reference_begin, alignment.reference_end = out.ref_begin, out.ref_end
        alignment.query_begin, alignment.query_end = out.query_begin, out.query_end
        alignment.ref_end_next_best = -1 if out.score else 0      # ssw.c:838-841 with maskLen < 15
        alignment.mismatches = out.mismatches
        alignment.cigar_string = cigar.value.decode()
        return True


def _scoring():
    return native.hs_scoring(StitchOptions.MATCH_PENALTY, StitchOptions.MISMATCH_PENALTY,
                             StitchOptions.GAP_PENALTY, StitchOptions.GAP_EXTEND_PENALTY)


def decode_region(positions, bases, rles):
    """Position dictionary + label decoding of small_chunk_stitch (Stitch.py:214-245) for the concatenated
    rows of one region's chunks, in the order the reference visits them."""
    positions = np.ascontiguousarray(positions, dtype=np.int64).reshape(-1, 3)
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1)
    rles = np.ascontiguousarray(rles, dtype=np.uint8).reshape(-1)
    if not (len(positions) == len(bases) == len(rles)):
        raise ValueError("positions, bases and rles must have one row per prediction")
    cap = int(rles.astype(np.int64).sum()) + 1
    out = ctypes.create_string_buffer(cap)
    n = native.check(native.load().hs_decode_region(
        positions.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), bases.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
        rles.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), len(bases), out, cap))
    return out.raw[:n].decode()


class Stitch:
    """Joins the sequences predicted for overlapping regions of one contig (Stitch.py:14-31)."""

    def __init__(self):
        self.last_warnings = (0, 0, 0)       # no alignment / no anchor / no overlap, from the last alignment_stitch
        self.last_alignments = 0

    @staticmethod
    def get_confident_positions(alignment):
        """Anchor of an alignment: start of the first run of >= OVERLAP_THRESHOLD aligned bases, (-1, -1) if
        there is none (Stitch.py:34-94)."""
        ref_pos, query_pos = ctypes.c_int32(), ctypes.c_int32()
        native.check(native.load().hs_anchor_from_cigar(alignment.cigar_string.encode(), alignment.reference_begin,
                                                       StitchOptions.OVERLAP_THRESHOLD, ctypes.byref(ref_pos),
                                                       ctypes.byref(query_pos)))
        return ref_pos.value, query_pos.value

    def alignment_stitch(self, sequence_chunks):
        """sequence_chunks: (contig, start, end, sequence) tuples of one contig -> one tuple (Stitch.py:96-193)."""
        lib = native.load()
        scoring = _scoring()
        handle = lib.hs_stitcher_create(ctypes.byref(scoring), StitchOptions.OVERLAP_THRESHOLD, StitchOptions.BASE_ERROR_RATE)
        if not handle:
            raise RuntimeError("hs_stitcher_create failed: " + lib.hs_last_error().decode())
        try:
            ordered = sorted(sequence_chunks, key=lambda element: (element[1], element[2]))
            contig = ordered[0][0]
            total = 0
            for _, start, end, sequence in ordered:
                raw = sequence.encode()
                total += len(raw) + 10
                native.check(lib.hs_stitcher_add(handle, int(start), int(end), raw, len(raw)))
            start, end, length = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
            warnings = (ctypes.c_int64 * 3)()
            alignments = ctypes.c_int64()
            native.check(lib.hs_stitcher_run(handle, ctypes.byref(start), ctypes.byref(end), ctypes.byref(length),
                                             warnings, ctypes.byref(alignments)))
            out = ctypes.create_string_buffer(length.value + 1)
            native.check(lib.hs_stitcher_sequence(handle, out, length.value + 1))
        finally:
            lib.hs_stitcher_destroy(handle)
        self.last_warnings, self.last_alignments = tuple(warnings), alignments.value
        if any(self.last_warnings):
            sys.stderr.write(TextColor.YELLOW + "WARNING: " + str(contig) + " " + str(start.value) + "-" + str(end.value)
                             + ": NO ALIGNMENT FOUND x" + str(warnings[0]) + ", NO OVERLAPS IN ALIGNMENT x" + str(warnings[1])
                             + ", NO OVERLAP IN CHUNKS x" + str(warnings[2]) + "\n" + TextColor.END)
        return contig, start.value, end.value, out.raw[:length.value].decode()

    def small_chunk_stitch(self, contig, small_chunk_keys):
        """Regions of one contig -> one stitched piece (Stitch.py:195-254).  Inside a region the images share
        one coordinate system, so their predictions are merged by position; regions are then joined by
        alignment_stitch."""
        name_sequence_tuples = list()
        for contig_name, file_name, chunk_name, contig_start, contig_end in small_chunk_keys:
            positions, bases, rles = [], [], []
            native_reader = region_reader(file_name)
            if native_reader is not None:
                # the region's rows in one library call (no interpreter lock: the worker threads read in parallel)
                try:
                    rows = native_reader.read_prediction_region(contig, chunk_name)
                    sequence = decode_region(*rows) if len(rows[1]) else ''
                    name_sequence_tuples.append((contig, contig_start, contig_end, sequence))
                    continue
                except (_feed_native.Unsupported, IOError):
                    pass                                              # a packed file, or one outside the library's subset
            with open_predictions(file_name) as hdf5_file:            # one open per region instead of one per image
                if 'predictions' in hdf5_file:
                    region = hdf5_file['predictions'][contig][chunk_name]
                    for chunk in sorted(set(region.keys()) - {'contig_start', 'contig_end'}):
                        bases.append(np.asarray(region[chunk]['bases'][()]).reshape(-1))
                        rles.append(np.asarray(region[chunk]['rles'][()]).reshape(-1))
                        positions.append(np.asarray(region[chunk]['position'][()], dtype=np.int64).reshape(-1, 3))
            if positions:
                sequence = decode_region(np.concatenate(positions), np.concatenate(bases), np.concatenate(rles))
            else:
                sequence = ''
            name_sequence_tuples.append((contig, contig_start, contig_end, sequence))
        return self.alignment_stitch(name_sequence_tuples)

    def create_consensus_sequence(self, contig, sequence_chunk_keys, threads):
        """All regions of a contig -> its consensus sequence (Stitch.py:256-301): groups of consecutive regions
        are stitched by worker threads, the group results are stitched once more."""
        sequence_chunk_key_list = [(contig, hdf5_file, chunk_key, int(st), int(end))
                                   for hdf5_file, chunk_key, st, end in sequence_chunk_keys]
        sequence_chunk_key_list = sorted(sequence_chunk_key_list, key=lambda element: (element[3], element[4]))
        if not sequence_chunk_key_list:
            return ''
        group = max(StitchOptions.MIN_SEQUENCE_REQUIRED_FOR_MULTITHREADING, int(len(sequence_chunk_key_list) / threads) + 1)
        file_chunks = list(FileManager.chunks(sequence_chunk_key_list, group))
        sequence_chunks = list()
        if threads <= 1 or len(file_chunks) == 1:
            for file_chunk in file_chunks:
                sequence_chunks.append(self.small_chunk_stitch(contig, file_chunk))
        else:
            # threads, not the reference's processes: the library calls run without the interpreter lock and
            # nothing has to be pickled
            with concurrent.futures.ThreadPoolExecutor(max_workers=threads) as executor:
                futures = [executor.submit(Stitch().small_chunk_stitch, contig, file_chunk) for file_chunk in file_chunks]
                for fut in concurrent.futures.as_completed(futures):
                    if fut.exception() is None:
                        sequence_chunks.append(fut.result())
                    else:
                        sys.stderr.write("ERROR: " + str(fut.exception()) + "\n")
        sequence_chunks = sorted(sequence_chunks, key=lambda element: (element[1], element[2]))
        contig, contig_start, contig_end, sequence = self.alignment_stitch(sequence_chunks)
        return sequence
