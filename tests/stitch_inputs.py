"""Deterministic inputs for the stitch tests, shared by tests/golden/make_golden_stitch.py (which runs the
reference's own Stitch class on them) and tests/test_stitch.py (which runs helen_b200's)."""
import random

import numpy as np

BASES = "ACGT"


def random_sequence(rng, n, alphabet=BASES):
    return ''.join(rng.choice(alphabet) for _ in range(n))


def with_errors(rng, seq, rate, alphabet=BASES):
    """Substitutions, deletions and insertions (lengths 1..12) at `rate` per base."""
    out, i = [], 0
    while i < len(seq):
        x = rng.random()
        if x < rate / 3:
            out.append(rng.choice(alphabet))
            i += 1
        elif x < 2 * rate / 3:
            i += rng.choice([1, 1, 1, 2, 5, 12])
        elif x < rate:
            out.append(random_sequence(rng, rng.choice([1, 1, 1, 2, 5, 12]), alphabet))
        else:
            out.append(seq[i])
            i += 1
    return ''.join(out)


def aligner_pairs(seed=11, count=400):
    """(reference, query) pairs: overlap-like pairs at several error rates, unrelated pairs, low-complexity
    repeats, sequences with N, lengths around the 16- and 8-lane segment boundaries and around the
    8-bit -> 16-bit kernel switch (score 249)."""
    rng = random.Random(seed)
    pairs = [("ACGT", "ACGT"), ("A", "A"), ("A", "C"), ("NNNN", "NNNN"), ("ACGTACGTAC", "TTTT"),
             ("A" * 70, "A" * 61), ("AC" * 40, "CA" * 33), ("ACGTTGCA" * 9, "ACGTTGCA" * 8)]
    while len(pairs) < count:
        length = rng.choice([1, 2, 5, 15, 16, 17, 31, 33, 40, 62, 63, 64, 80, 120, 200, 300, 500])
        alphabet = rng.choice([BASES, BASES, BASES, "ACGTN", "AC"])
        a = random_sequence(rng, length, alphabet)
        mode = rng.random()
        if mode < 0.6:
            b = with_errors(rng, a, rng.choice([0.02, 0.05, 0.15, 0.3]), alphabet)
            b = b[rng.randrange(0, max(1, length // 2)):] + random_sequence(rng, rng.randrange(0, 20), alphabet)
        elif mode < 0.8:
            b = random_sequence(rng, rng.choice([1, 3, 10, 50, 200]), alphabet)
        else:
            unit = random_sequence(rng, rng.choice([1, 2, 3, 7]))
            a = (unit * (length // len(unit) + 1))[:length]
            b = list(a)
            for _ in range(rng.randrange(0, 4)):
                p, k = rng.randrange(0, length), rng.choice([3, 7, 8, 10, 16, 25])
                b[p:p + k] = random_sequence(rng, rng.choice([k, k, 0, k + 3, 2 * k]))
            b = ''.join(b)
        pairs.append((a, b or "A"))
    return pairs


def region_pieces(seed, regions=8, region_len=600, overlap=60, rate=0.02, special=()):
    """(contig, start, end, sequence) pieces as alignment_stitch receives them: noisy copies of overlapping
    slices of one truth sequence.  `special`: kinds of trouble injected at random pieces."""
    rng = random.Random(seed)
    truth = random_sequence(rng, regions * region_len + overlap)
    pieces = []
    for r in range(regions):
        start = r * region_len
        end = min(len(truth), start + region_len + overlap)
        pieces.append(["contig_%d" % seed, start, end, with_errors(rng, truth[start:end], rate)])
    for kind in special:
        k = rng.randrange(1, regions)
        if kind == "gap":                 # the piece starts after the running end: no overlap in chunks
            pieces[k][1] += overlap + 25
        elif kind == "short":             # too short to be worth adding
            pieces[k][3] = pieces[k][3][:7]
        elif kind == "empty":
            pieces[k][3] = ""
        elif kind == "garbage":           # overlapping head unrelated to the running tail: no anchor / no alignment
            pieces[k][3] = random_sequence(rng, len(pieces[k][3]), "AC" if rng.random() < 0.5 else BASES)
        elif kind == "ns":
            pieces[k][3] = "N" * len(pieces[k][3])
        elif kind == "contained":         # a piece inside its predecessor
            pieces[k][2] = pieces[k - 1][2] - 5
        elif kind == "duplicate":
            pieces.insert(k, list(pieces[k]))
        elif kind == "long_overlap":      # the overlap exceeds the incoming piece
            pieces[k][1] = max(0, pieces[k][1] - region_len)
    rng.shuffle(pieces)
    return [tuple(p) for p in pieces]


def prediction_records(seed, regions=5, images_per_region=3, window=1000, image_overlap=200, contig="chr_t"):
    """Records as predict_gpu.py:176-179 hands them to DataStore.write_prediction: per region a few images whose
    position rows overlap (same coordinates, possibly different predictions), insert columns (index > 0),
    and (-1, -1, -1) padding rows at the end of the last image.  Adjacent regions overlap by 40 positions and
    agree there except for the 3 % of predictions flipped per image."""
    rng = random.Random(seed)
    np_rng = np.random.default_rng(seed)
    region_span = images_per_region * (window - image_overlap) // 2
    region_step = region_span - 40
    # the coordinate rows of the whole contig: (ref position, insert index, split index), one truth per row
    rows = []
    for pos in range(region_step * (regions - 1) + region_span):
        rows.append((pos, 0, 0))
        for ins in range(1, 1 + (rng.random() < 0.04) * rng.choice([1, 1, 2, 3])):
            rows.append((pos, ins, 0))
        if rng.random() < 0.02:
            rows.append((pos, 0, 1))              # a run split over two columns
    rows = np.array(rows, dtype=np.int64)
    # mostly runs of one base, so that sequence length stays close to the position span (as the overlap arithmetic
    # of alignment_stitch assumes) and neighbouring regions usually find an anchor
    truth_base = np.where(np_rng.random(len(rows)) < 0.06, 0, np_rng.integers(1, 5, len(rows))).astype(np.uint8)
    truth_rle = np.where(np_rng.random(len(rows)) < 0.92, 1, np_rng.integers(2, 7, len(rows)))
    truth_rle = np.where(truth_base == 0, 0, truth_rle).astype(np.uint8)
    records = []
    for r in range(regions):
        region_start = r * region_step
        region_end = region_start + region_span
        inside = np.flatnonzero((rows[:, 0] >= region_start) & (rows[:, 0] < region_end))
        first, count = int(inside[0]), len(inside)
        step = max(1, (count - window) // max(1, images_per_region - 1)) if count > window else count
        for c in range(images_per_region):
            lo = min(c * step, max(0, count - window))
            hi = min(count, lo + window)
            position = np.full((window, 3), -1, dtype=np.int64)
            position[:hi - lo] = rows[first + lo:first + hi]
            bases = np.zeros(window, dtype=np.uint8)
            rles = np.zeros(window, dtype=np.uint8)
            bases[:hi - lo], rles[:hi - lo] = truth_base[first + lo:first + hi], truth_rle[first + lo:first + hi]
            flip = np_rng.random(hi - lo) < 0.03           # overlapping images may disagree; the first one kept wins
            bases[:hi - lo][flip] = np_rng.integers(0, 5, int(flip.sum()))
            rles[:hi - lo][flip] = np_rng.integers(0, 7, int(flip.sum()))
            records.append((contig, region_start, region_end, c, position, bases, rles))
    return records
