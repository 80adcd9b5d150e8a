"""Mints tests/golden/stitch_cases.json from the reference's OWN stitch code, run in the build container:

* `HELEN.Aligner` = the reference's pybind module, compiled from /root/reference's ssw.c / ssw_cpp.cpp /
  pybind_api.cpp by oracle/ssw_ref/build_ref.py (-> oracle/_ref/helen/build/HELEN*.so);
* `Stitch` = helen/modules/python/Stitch.py imported unmodified from /root/reference.

Shims (none touches the stitch logic): `helen.build` is registered by hand because the reference expects its
cmake build to drop HELEN.so into helen/build/; h5py (absent from this image) is tests/fake_h5's in-memory
stand-in; `np.int` (removed in numpy 1.24, used at Stitch.py:223-224) is aliased to int.

    python tests/golden/make_golden_stitch.py
"""
import glob
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

from oracle.ssw_ref import build_ref  # noqa: E402

build_ref.build()
ext = glob.glob(os.path.join(ROOT, "oracle", "_ref", "helen", "build", "HELEN*.so"))[0]
spec = importlib.util.spec_from_file_location("HELEN", ext)
HELEN = importlib.util.module_from_spec(spec)
spec.loader.exec_module(HELEN)
build_pkg = types.ModuleType("helen.build")
build_pkg.HELEN = HELEN
import helen  # noqa: E402  (the reference package)
sys.modules["helen.build"] = build_pkg
helen.build = build_pkg

import fake_h5  # noqa: E402
h5py = types.ModuleType("h5py")
h5py.File = fake_h5.FakeFile
sys.modules["h5py"] = h5py
if not hasattr(np, "int"):
    np.int = int

from helen.modules.python.Stitch import Stitch  # noqa: E402
from helen.modules.python.Options import StitchOptions  # noqa: E402
import stitch_inputs  # noqa: E402

import helen_b200.hdf5 as hb_hdf5  # noqa: E402  (only to WRITE the prediction files the reference then reads)
from helen_b200.DataStore import DataStore  # noqa: E402

hb_hdf5.open_file = fake_h5.open_file

PIECE_CASES = [
    dict(seed=1), dict(seed=2, rate=0.08), dict(seed=3, special=["gap"]), dict(seed=4, special=["short", "garbage"]),
    dict(seed=5, special=["empty"]), dict(seed=6, special=["ns"]), dict(seed=7, special=["contained", "duplicate"]),
    dict(seed=8, rate=0.3), dict(seed=9, special=["long_overlap"]), dict(seed=10, regions=1),
    dict(seed=11, regions=14, region_len=250, overlap=120, rate=0.05, special=["garbage", "gap", "short"]),
    dict(seed=12, regions=6, region_len=40, overlap=12, rate=0.1),
]
RECORD_CASES = [dict(seed=21), dict(seed=22, regions=3, images_per_region=4), dict(seed=23, regions=7, images_per_region=2)]


def main():
    out = {"options": {k: getattr(StitchOptions, k) for k in ("MATCH_PENALTY", "MISMATCH_PENALTY", "GAP_PENALTY",
                                                               "GAP_EXTEND_PENALTY", "OVERLAP_THRESHOLD", "BASE_ERROR_RATE")},
           "aligner": [], "alignment_stitch": [], "small_chunk_stitch": [], "create_consensus_sequence": []}

    aligner = HELEN.Aligner(StitchOptions.MATCH_PENALTY, StitchOptions.MISMATCH_PENALTY, StitchOptions.GAP_PENALTY,
                            StitchOptions.GAP_EXTEND_PENALTY)
    flt = HELEN.Filter()
    for ref, query in stitch_inputs.aligner_pairs():
        al = HELEN.Alignment()
        aligner.SetReferenceSequence(ref, len(ref))
        aligner.Align_cpp(query, flt, al, 0)
        row = {"score": al.best_score}
        if al.best_score:
            row.update(ref_begin=al.reference_begin, ref_end=al.reference_end, query_begin=al.query_begin,
                       query_end=al.query_end, mismatches=al.mismatches, cigar=al.cigar_string,
                       anchor=list(Stitch.get_confident_positions(al)))
        out["aligner"].append(row)

    for kwargs in PIECE_CASES:
        pieces = stitch_inputs.region_pieces(**kwargs)
        contig, start, end, seq = Stitch().alignment_stitch(pieces)
        out["alignment_stitch"].append({"kwargs": kwargs, "contig": contig, "start": start, "end": end, "sequence": seq})

    for kwargs in RECORD_CASES:
        fake_h5.reset()
        path = "/golden/pred_%d.hdf" % kwargs["seed"]
        store = DataStore(path, mode='w')
        records = stitch_inputs.prediction_records(**kwargs)
        for contig, start, end, chunk_id, position, bases, rles in records:
            store.write_prediction(contig, start, end, chunk_id, position, bases, rles)
        store.close()
        contig = records[0][0]
        regions = sorted({(contig, path, "%s-%d-%d" % (contig, s, e), s, e) for _, s, e, *_ in records}, key=lambda k: (k[3], k[4]))
        got = Stitch().small_chunk_stitch(contig, regions)
        out["small_chunk_stitch"].append({"kwargs": kwargs, "contig": got[0], "start": int(got[1]), "end": int(got[2]),
                                          "sequence": got[3]})
        keys = [(path, name, s, e) for _, path, name, s, e in regions]
        for threads in (1, 3):
            seq = Stitch().create_consensus_sequence(contig, keys, threads)
            out["create_consensus_sequence"].append({"kwargs": kwargs, "threads": threads, "sequence": seq})

    with open(os.path.join(HERE, "stitch_cases.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
