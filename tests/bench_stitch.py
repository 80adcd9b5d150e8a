"""CPU benchmark of the stitch library: local alignments per second and whole
alignment_stitch runs, next to the reference's own Smith-Waterman compiled into oracle/_ref (and, where
/root/reference is present, the reference's Python Stitch class over its pybind module).  Single core.

    python tests/bench_stitch.py [--json profiles/r01_stitch_bench.json]
"""
import argparse
import ctypes
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))        # lives under tests/ because it executes oracle/ (the reference, as baseline)

import stitch_inputs  # noqa: E402
from helen_b200 import _stitch_native as native  # noqa: E402
from helen_b200 import build as hb_build  # noqa: E402


def overlap_pairs(rng, count, length, rate):
    pairs = []
    for _ in range(count):
        truth = stitch_inputs.random_sequence(rng, 2 * length)
        left = stitch_inputs.with_errors(rng, truth[:length + length // 2], rate)[-length:]
        right = stitch_inputs.with_errors(rng, truth[length // 2:], rate)[:length]
        pairs.append((left, right))
    return pairs


def time_aligner(fn, pairs, min_seconds=1.0):
    fn(*pairs[0])
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < min_seconds:
        for ref, query in pairs:
            fn(ref, query)
        done += len(pairs)
    return done / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    hb_build.build_stitch()
    lib = native.load()
    scoring = native.hs_scoring(4, 6, 8, 2)

    def mine(ref, query):
        out = native.hs_alignment()
        cigar = ctypes.create_string_buffer(4096)
        native.check(lib.hs_ssw_align(ref.encode(), len(ref), query.encode(), len(query), ctypes.byref(scoring),
                                      ctypes.byref(out), cigar, 4096))
        return out.score

    from oracle import ssw_ref
    have_ref = ssw_ref.load() is not None
    rng = random.Random(0)
    result = {"host": os.uname().nodename, "cores_used": 1, "aligner": [], "alignment_stitch": []}
    for length in (50, 100, 200, 400):
        pairs = overlap_pairs(rng, 64, length, 0.02)
        row = {"overlap_bases": length, "helen_b200_alignments_per_s": round(time_aligner(mine, pairs))}
        if have_ref:
            row["reference_alignments_per_s"] = round(time_aligner(lambda r, q: ssw_ref.align(r, q), pairs))
        result["aligner"].append(row)
        print(row)

    from helen_b200.Stitch import Stitch
    ref_stitch = None
    if os.path.isdir("/root/reference/helen"):
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
            import make_golden_stitch  # noqa: F401  (installs the shims and imports the reference's Stitch)
            from helen.modules.python.Stitch import Stitch as ref_stitch
        except Exception as exc:       # the reference's Python side is optional here
            print("reference Stitch class unavailable:", exc)
    for regions in (100, 1000):
        pieces = stitch_inputs.region_pieces(seed=3, regions=regions, region_len=1000, overlap=100, rate=0.01)
        t0 = time.perf_counter()
        got = Stitch().alignment_stitch(pieces)
        mine_s = time.perf_counter() - t0
        row = {"regions": regions, "bases": len(got[3]), "helen_b200_s": round(mine_s, 4)}
        if ref_stitch is not None:
            devnull, saved = open(os.devnull, "w"), sys.stderr
            sys.stderr = devnull
            try:
                t0 = time.perf_counter()
                want = ref_stitch().alignment_stitch(pieces)
                row["reference_s"] = round(time.perf_counter() - t0, 4)
            finally:
                sys.stderr = saved
            row["identical"] = tuple(want) == tuple(got)
        result["alignment_stitch"].append(row)
        print(row)
    # regions -> sequences: position merge + label decoding + joins, from in-memory prediction files
    import fake_h5
    import helen_b200.hdf5 as hb_hdf5
    from helen_b200.DataStore import DataStore
    hb_hdf5.open_file = fake_h5.open_file
    fake_h5.reset()
    records = stitch_inputs.prediction_records(seed=4, regions=60, images_per_region=3)
    store = DataStore("/bench/pred.hdf", mode='w')
    for contig, start, end, chunk_id, position, bases, rles in records:
        store.write_prediction(contig, start, end, chunk_id, position, bases, rles)
    store.close()
    contig = records[0][0]
    regions = sorted({(contig, "/bench/pred.hdf", "%s-%d-%d" % (contig, s, e), s, e) for _, s, e, *_ in records},
                     key=lambda k: (k[3], k[4]))
    t0 = time.perf_counter()
    got = Stitch().small_chunk_stitch(contig, regions)
    row = {"regions": len(regions), "images": len(records), "bases": len(got[3]), "helen_b200_s": round(time.perf_counter() - t0, 4)}
    if ref_stitch is not None:
        devnull, saved = open(os.devnull, "w"), sys.stderr
        sys.stderr = devnull
        try:
            t0 = time.perf_counter()
            want = ref_stitch().small_chunk_stitch(contig, regions)
            row["reference_s"] = round(time.perf_counter() - t0, 4)
        finally:
            sys.stderr = saved
        row["identical"] = tuple(want) == tuple(got)
    result["small_chunk_stitch"] = row
    print(row)
    if args.json:
        with open(args.json, "w") as fh:
            json.dump(result, fh, indent=1)


if __name__ == "__main__":
    main()
