"""Stitch library against the reference: golden vectors minted from the
reference's own Stitch class and pybind aligner (tests/golden/make_golden_stitch.py), a differential run
against the reference's Smith-Waterman compiled into oracle/_ref, and the host-side behaviour around it.
All CPU: the library has no GPU part.  Bar: identical integers and identical strings."""
import ctypes
import json
import os
import random
import re

import numpy as np
import pytest

import fake_h5
import stitch_inputs
from helen_b200 import _stitch_native as native
from helen_b200 import build as hb_build
from helen_b200.options import StitchOptions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    hb_build.build_stitch()
    return native.load()


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "stitch_cases.json")) as fh:
        return json.load(fh)


@pytest.fixture()
def in_memory_h5(monkeypatch):
    import helen_b200.hdf5 as hb_hdf5
    fake_h5.reset()
    from helen_b200.DataStore import forget_packed_views
    forget_packed_views()
    monkeypatch.setattr(hb_hdf5, "open_file", fake_h5.open_file)
    yield
    fake_h5.reset()
    forget_packed_views()


def my_align(lib, ref, query, scoring=(4, 6, 8, 2)):
    out = native.hs_alignment()
    sc = native.hs_scoring(*scoring)
    cigar = ctypes.create_string_buffer(16 * (len(ref) + len(query)) + 64)
    native.check(lib.hs_ssw_align(ref.encode(), len(ref), query.encode(), len(query), ctypes.byref(sc),
                                  ctypes.byref(out), cigar, len(cigar)))
    if out.score == 0:
        return dict(score=0)
    return dict(score=out.score, ref_begin=out.ref_begin, ref_end=out.ref_end, query_begin=out.query_begin,
                query_end=out.query_end, mismatches=out.mismatches, cigar=cigar.value.decode())


def test_header_and_binding_agree(lib):
    text = open(os.path.join(ROOT, "include", "helen_stitch.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(hs_[a-z_0-9]+)\s*\(", text)))
    assert declared and sorted(native.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/helen_stitch.h but not exported"
    assert lib.hs_abi_version() == native.HS_ABI_VERSION


def test_options_match_reference(golden):
    for key, value in golden["options"].items():
        assert getattr(StitchOptions, key) == value


def test_aligner_matches_reference_golden(lib, golden):
    """400 pairs through the reference's HELEN.Aligner (Stitch.py:110-135): score, ends, begins, mismatch count,
    cigar string and the anchor get_confident_positions derives from it."""
    from helen_b200.Stitch import Aligner, Alignment, Filter, Stitch
    pairs = stitch_inputs.aligner_pairs()
    assert len(pairs) == len(golden["aligner"])
    aligner = Aligner(StitchOptions.MATCH_PENALTY, StitchOptions.MISMATCH_PENALTY, StitchOptions.GAP_PENALTY,
                      StitchOptions.GAP_EXTEND_PENALTY)
    kernels = set()
    for (ref, query), want in zip(pairs, golden["aligner"]):
        al = Alignment()
        aligner.SetReferenceSequence(ref, len(ref))
        aligner.Align_cpp(query, Filter(), al, 0)
        assert al.best_score == want["score"], (ref, query)
        if want["score"]:
            got = dict(ref_begin=al.reference_begin, ref_end=al.reference_end, query_begin=al.query_begin,
                       query_end=al.query_end, mismatches=al.mismatches, cigar=al.cigar_string,
                       anchor=list(Stitch.get_confident_positions(al)))
            assert got == {k: want[k] for k in got}, (ref, query)
            kernels.add(want["score"] >= 249)
    assert kernels == {False, True}, "the pairs must exercise both of the reference's kernels"


def test_alignment_stitch_matches_reference_golden(lib, golden):
    from helen_b200.Stitch import Stitch
    seen = [0, 0, 0]
    for want in golden["alignment_stitch"]:
        stitcher = Stitch()
        got = stitcher.alignment_stitch(stitch_inputs.region_pieces(**want["kwargs"]))
        assert got == (want["contig"], want["start"], want["end"], want["sequence"]), want["kwargs"]
        seen = [a + b for a, b in zip(seen, stitcher.last_warnings)]
    assert all(seen), f"every warning branch of Stitch.py:138-190 must be exercised, saw {seen}"


def _write_records(path, records):
    from helen_b200.DataStore import DataStore
    store = DataStore(path, mode='w')
    for contig, start, end, chunk_id, position, bases, rles in records:
        store.write_prediction(contig, start, end, chunk_id, position, bases, rles)
    store.close()
    contig = records[0][0]
    regions = sorted({(contig, path, "%s-%d-%d" % (contig, s, e), s, e) for _, s, e, *_ in records}, key=lambda k: (k[3], k[4]))
    return contig, regions


def test_small_chunk_stitch_matches_reference_golden(lib, golden, in_memory_h5):
    from helen_b200.Stitch import Stitch
    for want in golden["small_chunk_stitch"]:
        contig, regions = _write_records("/t/pred.hdf", stitch_inputs.prediction_records(**want["kwargs"]))
        got = Stitch().small_chunk_stitch(contig, regions)
        assert got == (want["contig"], want["start"], want["end"], want["sequence"]), want["kwargs"]
        fake_h5.reset()


def test_create_consensus_sequence_matches_reference_golden(lib, golden, in_memory_h5):
    from helen_b200.Stitch import Stitch
    for want in golden["create_consensus_sequence"]:
        contig, regions = _write_records("/t/pred.hdf", stitch_inputs.prediction_records(**want["kwargs"]))
        keys = [(path, name, s, e) for _, path, name, s, e in regions]
        assert Stitch().create_consensus_sequence(contig, keys, want["threads"]) == want["sequence"], want
        fake_h5.reset()


def test_differential_against_compiled_reference(lib):
    """12,000 random pairs, four scorings, against oracle/_ref/libssw_ref.so (the reference's ssw.c + ssw_cpp.cpp)."""
    from oracle import ssw_ref
    if ssw_ref.load() is None:
        pytest.skip("oracle/_ref/libssw_ref.so not built and /root/reference not present")
    rng = random.Random(5)
    scorings = [(4, 6, 8, 2), (2, 2, 3, 1), (1, 4, 6, 1), (3, 9, 4, 1)]
    checked = 0
    for round_ in range(30):
        for ref, query in stitch_inputs.aligner_pairs(seed=100 + round_, count=400):
            scoring = scorings[0] if rng.random() < 0.6 else rng.choice(scorings)
            want = ssw_ref.align(ref, query, *scoring)
            got = my_align(lib, ref, query, scoring)
            assert got == want, (scoring, ref, query)
            checked += want["score"] > 0
    assert checked > 10000


def test_every_sweep_variant_agrees_with_the_reference(lib):
    """The column sweep has three code paths: AVX2 (16 rows per instruction, when the CPU has it), SSE2 (8 rows;
    forced by HS_NO_AVX2=1, read once per process, hence the subprocess) and plain 32-bit cells for inputs whose
    scores could leave 16 bits.  Each against the compiled reference."""
    import subprocess
    import sys
    from oracle import ssw_ref
    if ssw_ref.load() is None:
        pytest.skip("oracle/_ref/libssw_ref.so not built and /root/reference not present")
    script = (
        "import sys; sys.path[:0] = [%r, %r]\n"
        "import ctypes, stitch_inputs\n"
        "from helen_b200 import _stitch_native as native\n"
        "from oracle import ssw_ref\n"
        "lib = native.load(); sc = native.hs_scoring(4, 6, 8, 2); n = 0\n"
        "for ref, query in stitch_inputs.aligner_pairs(seed=77, count=1500):\n"
        "    out = native.hs_alignment(); cigar = ctypes.create_string_buffer(16 * (len(ref) + len(query)) + 64)\n"
        "    native.check(lib.hs_ssw_align(ref.encode(), len(ref), query.encode(), len(query), ctypes.byref(sc), ctypes.byref(out), cigar, len(cigar)))\n"
        "    want = ssw_ref.align(ref, query)\n"
        "    if want['score'] == 0:\n"
        "        assert out.score == 0; continue\n"
        "    got = dict(score=out.score, ref_begin=out.ref_begin, ref_end=out.ref_end, query_begin=out.query_begin, query_end=out.query_end, mismatches=out.mismatches, cigar=cigar.value.decode())\n"
        "    assert got == want, (ref, query); n += 1\n"
        "print('checked', n)\n") % (ROOT, os.path.join(ROOT, "tests"))
    for env_extra in ({"HS_NO_AVX2": "1"}, {}):
        proc = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env={**os.environ, **env_extra})
        assert proc.returncode == 0 and "checked" in proc.stdout, proc.stdout + proc.stderr
    # 32-bit cells: 7,400 matching bases (score 29,000+, still inside the reference's 16-bit kernel)
    rng = random.Random(3)
    ref = stitch_inputs.random_sequence(rng, 7600)
    query = stitch_inputs.with_errors(rng, ref[100:7500], 0.002)
    assert my_align(lib, ref, query) == ssw_ref.align(ref, query)
    with pytest.raises(RuntimeError):                   # 36,000 > 32,767: beyond the reference's kernel, refused
        my_align(lib, "ACGT" * 2250, "ACGT" * 2250)


def test_property_based_against_compiled_reference(lib):
    """hypothesis-generated pairs (short, repetitive, N-rich: where ties between equally scoring alignments are
    most frequent) and scorings with gap_open > gap_extend."""
    from hypothesis import given, settings, strategies as st
    from oracle import ssw_ref
    if ssw_ref.load() is None:
        pytest.skip("oracle/_ref/libssw_ref.so not built and /root/reference not present")
    sequence = st.text(alphabet="ACGTN", min_size=1, max_size=90)
    tandem = st.builds(lambda unit, n, tail: unit * n + tail, st.text(alphabet="ACGT", min_size=1, max_size=4),
                       st.integers(1, 30), st.text(alphabet="ACGT", max_size=6))
    scoring = st.tuples(st.integers(1, 6), st.integers(0, 9), st.integers(2, 12), st.integers(0, 4)).filter(lambda s: s[2] > s[3])

    @settings(max_examples=600, deadline=None, derandomize=True)
    @given(st.one_of(sequence, tandem), st.one_of(sequence, tandem), scoring)
    def check(ref, query, sc):
        assert my_align(lib, ref, query, sc) == ssw_ref.align(ref, query, *sc), (ref, query, sc)
    check()


def test_alignment_properties_at_length(lib):
    """Size-independent properties on inputs longer than any fixture (2,000-base overlaps): the cigar consumes
    exactly the aligned spans, its score re-derived from the cigar equals the reported score, an exact copy aligns
    end to end."""
    rng = random.Random(9)
    match, mismatch, gap_open, gap_extend = 4, 6, 8, 2
    for _ in range(6):
        ref = stitch_inputs.random_sequence(rng, 2000)
        query = stitch_inputs.with_errors(rng, ref[300:], 0.03)
        al = my_align(lib, ref, query)
        ops = [(int(n), op) for n, op in re.findall(r"(\d+)([=XIDS])", al["cigar"])]
        assert "".join(f"{n}{op}" for n, op in ops) == al["cigar"]
        assert sum(n for n, op in ops if op in "=XD") == al["ref_end"] - al["ref_begin"] + 1
        assert sum(n for n, op in ops if op in "=XI") == al["query_end"] - al["query_begin"] + 1
        assert sum(n for n, op in ops if op in "=XIS") == len(query)
        score = sum({"=": match * n, "X": -mismatch * n, "S": 0}.get(op, -(gap_open + gap_extend * (n - 1))) for n, op in ops)
        assert score == al["score"]
        assert al["mismatches"] == sum(n for n, op in ops if op in "XID")
    exact = my_align(lib, ref, ref[100:1500])
    assert exact == dict(score=4 * 1400, ref_begin=100, ref_end=1499, query_begin=0, query_end=1399, mismatches=0, cigar="1400=")


def test_edge_cases_and_errors(lib):
    from helen_b200.Stitch import Aligner, Alignment, Filter, Stitch, decode_region
    aligner, al = Aligner(4, 6, 8, 2), Alignment()
    aligner.SetReferenceSequence("", 0)
    assert aligner.Align_cpp("ACGT", Filter(), al, 0) is False and al.best_score == 0        # ssw_cpp.cpp:324-326
    aligner.SetReferenceSequence("ACGT", 4)
    assert aligner.Align_cpp("", Filter(), al, 0) is False and al.best_score == 0
    assert aligner.Align_cpp("NNNN", Filter(), al, 0) is True and al.best_score == 0 and al.cigar_string == ""
    assert my_align(lib, "acgtu", "ACGTA")["cigar"] == "5="                                  # case-insensitive, U reads as A
    with pytest.raises(ValueError):
        my_align(lib, "ACGT", "ACGT", (4, 6, 2, 2))          # gap_open must exceed gap_extend
    with pytest.raises(ValueError):
        Aligner(4, 6, 8, 2).Align_cpp("A", Filter(True, False, 0, 32767), al, 0)
    # anchors: '=' and 'X' runs merge; S and I advance the query, D the reference
    al.cigar_string, al.reference_begin = "3S4=2X3=1I9=", 5
    assert Stitch.get_confident_positions(al) == (5, 3)
    al.cigar_string = "3S4=1D3=1I9="
    assert Stitch.get_confident_positions(al) == (5 + 4 + 1 + 3, 3 + 4 + 3 + 1)
    al.cigar_string = "7=1I7=1D7="
    assert Stitch.get_confident_positions(al) == (-1, -1)
    al.cigar_string = "7=2N9="
    with pytest.raises(ValueError):
        Stitch.get_confident_positions(al)
    # decode: first prediction of a key wins, keys sorted, negative rows skipped, label 0 is a gap
    positions = np.array([[5, 0, 0], [3, 0, 0], [5, 0, 0], [-1, -1, -1], [3, 1, 0], [4, 0, 0], [3, 0, 1]])
    bases = np.array([1, 2, 3, 4, 4, 0, 3], dtype=np.uint8)
    rles = np.array([2, 1, 6, 6, 3, 5, 2], dtype=np.uint8)
    assert decode_region(positions, bases, rles) == "C" + "GG" + "TTT" + "" + "AA"
    assert decode_region(np.zeros((0, 3)), np.zeros(0), np.zeros(0)) == ""
    with pytest.raises(ValueError):
        decode_region(positions, np.full(7, 9, dtype=np.uint8), rles)
    assert Stitch().alignment_stitch([("c", 0, 5, "")]) == ("c", 0, 5, "")


def test_perform_stitch_writes_fasta(lib, in_memory_h5, tmp_path, monkeypatch):
    """Two prediction files (the per-rank files call_consensus writes), two contigs -> one FASTA record each,
    contigs in sorted order, same sequences as create_consensus_sequence."""
    import helen_b200.StitchInterface as iface
    from helen_b200.Stitch import Stitch
    rec_a = stitch_inputs.prediction_records(seed=31, regions=4, contig="chrB")
    rec_b = stitch_inputs.prediction_records(seed=32, regions=3, contig="chrA")
    paths = [str(tmp_path / "pred_0.hdf"), str(tmp_path / "pred_1.hdf")]
    contig_a, regions_a = _write_records(paths[0], rec_a[::2] + rec_b)       # chrB's images are split over both files
    from helen_b200.DataStore import DataStore
    store = DataStore(paths[1], mode='w')
    for contig, start, end, chunk_id, position, bases, rles in rec_a[1::2]:
        store.write_prediction(contig, start, end, chunk_id, position, bases, rles)
    store.close()
    monkeypatch.setattr(iface, "get_file_paths_from_directory", lambda directory: list(paths))
    out = iface.perform_stitch(str(tmp_path), str(tmp_path / "out"), "polished", 1)
    lines = open(out).read().splitlines()
    assert [l for l in lines if l.startswith(">")] == [">chrA", ">chrB"]
    fake_h5.reset()
    contig_b, regions_b = _write_records("/t/b.hdf", rec_b)
    want_a = Stitch().create_consensus_sequence("chrA", [(p, n, s, e) for _, p, n, s, e in regions_b], 1)
    assert lines[1] == want_a and len(lines[3]) > 1000


def test_polish_genome_ends_in_a_fasta(lib, in_memory_h5, tmp_path, monkeypatch):
    """polish_genome = call_consensus into predictions_<stamp>/, then stitch into <prefix>.fa
    (PolishInterface.py:49-105).  The GPU step is replaced by a writer of synthetic predictions here."""
    import helen_b200.PolishInterface as polish
    import helen_b200.StitchInterface as iface
    from helen_b200.Stitch import Stitch
    records = stitch_inputs.prediction_records(seed=41, regions=4, contig="chrP")
    written = {}

    def fake_call_consensus(image_dir, model_path, batch_size, num_workers, threads, output_dir, output_prefix, gpu_mode,
                            device_ids, callers):
        path = os.path.join(output_dir, output_prefix + "_0.hdf")
        written["contig"], written["regions"] = _write_records(path, records)
        written["path"] = path

    monkeypatch.setattr(polish, "call_consensus", fake_call_consensus)
    monkeypatch.setattr(iface, "get_file_paths_from_directory", lambda directory: [written["path"]])
    out_dir = str(tmp_path / "polish_out")
    prediction_dir = polish.polish_genome("imgs", "m.pkl", 8, 0, 2, out_dir, "HELEN_prediction", True, None, 1)
    assert os.path.basename(prediction_dir).startswith("predictions_") and os.path.dirname(prediction_dir) == out_dir
    lines = open(os.path.join(out_dir, "HELEN_prediction.fa")).read().splitlines()
    keys = [(p, n, s, e) for _, p, n, s, e in written["regions"]]
    assert lines == [">chrP", Stitch().create_consensus_sequence("chrP", keys, 1)] and len(lines[1]) > 2000


def test_packed_prediction_files_stitch_identically(lib, in_memory_h5, tmp_path, monkeypatch):
    """The packed schema (one HDF5 group per written batch instead of three datasets per image) read back through
    PackedPredictions gives the same FASTA as the reference schema, duplicates included."""
    import helen_b200.StitchInterface as iface
    from helen_b200.DataStore import DataStore, PackedPredictions, open_predictions
    records = stitch_inputs.prediction_records(seed=51, regions=5, contig="chrQ") + \
        stitch_inputs.prediction_records(seed=52, regions=3, contig="chrR")
    records.append(records[3])                                      # a repeated (region, chunk): the first one wins
    plain, packed = str(tmp_path / "plain_0.hdf"), str(tmp_path / "packed_0.hdf")
    _write_records(plain, records)
    store = DataStore(packed, mode='w', packed=True)
    for lo in range(0, len(records), 4):
        batch = records[lo:lo + 4]
        store.write_predictions([r[0] for r in batch], [r[1] for r in batch], [r[2] for r in batch], [r[3] for r in batch],
                                np.stack([r[4] for r in batch]), np.stack([r[5] for r in batch]), np.stack([r[6] for r in batch]))
    store.write_prediction(*records[0][:7])                        # single records go into a batch of one
    store.close()
    raw = fake_h5.open_file(packed)
    assert 'predictions' not in raw and len(raw['predictions_packed'].keys()) == (len(records) + 3) // 4 + 1
    view = open_predictions(packed)
    assert isinstance(view, PackedPredictions) and open_predictions(packed) is view
    assert sorted(view['predictions'].keys()) == ["chrQ", "chrR"]
    fasta = {}
    for name, path in (("plain", plain), ("packed", packed)):
        monkeypatch.setattr(iface, "get_file_paths_from_directory", lambda directory, path=path: [path])
        fasta[name] = open(iface.perform_stitch(str(tmp_path), str(tmp_path / ("out_" + name)), "p", 2)).read()
    assert fasta["packed"] == fasta["plain"] and fasta["plain"].count(">") == 2


def test_packed_mode_follows_the_environment(in_memory_h5, monkeypatch):
    from helen_b200.DataStore import DataStore
    monkeypatch.delenv("HELEN_B200_PACKED_PREDICTIONS", raising=False)
    assert DataStore("/t/a.hdf", "w").packed is False
    monkeypatch.setenv("HELEN_B200_PACKED_PREDICTIONS", "1")
    assert DataStore("/t/b.hdf", "w").packed is True and DataStore("/t/c.hdf", "w", packed=False).packed is False
