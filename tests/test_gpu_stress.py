"""Repeat-launch stress of every kernel variant of the tensor engine.

The roles of these kernels hand work to each other through mbarriers, counters in global memory and asynchronous
copies / tensor-memory stores; a missing wait shows up as a wrong result once in tens of launches, not in one.  (Two such
faults were found this way in round 2, see upload_whh in helen_b200/csrc/tensor_engine.cuh.)  Every variant is launched
REPS times on the same input and every launch must reproduce the fp32 engine within the stage tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle import random_state_dict

pytestmark = pytest.mark.gpu

REPS = int(os.environ.get("HB_STRESS_REPS", "120"))
ENV_KEYS = ("HB_WINDOWS_PER_CTA", "HB_NO_STACK", "HB_NO_PAIR", "HB_NO_PDL", "HB_NO_CHUNKLOOP", "HB_HEADS_WORKERS", "HB_NO_LIVE8",
            "HB_NO_PIXEL_JOBS", "HB_NO_PINGPONG", "HB_GATE_WARPS", "HB_NO_COOPERATIVE", "HB_NO_LOOP_PINGPONG", "HB_PIXELS_FIRST")
VARIANTS = {
    "product": {},
    "chunkloop_16_gate_warps": {"HB_GATE_WARPS": "16"},
    "chunkloop_tile16": {"HB_WINDOWS_PER_CTA": "16"},
    "chunkloop_tile16_one_tile": {"HB_WINDOWS_PER_CTA": "16", "HB_NO_LOOP_PINGPONG": "1"},
    "chunkloop_pixel_jobs_first": {"HB_PIXELS_FIRST": "1"},
    "chunkloop_tile16_16_gate_warps": {"HB_WINDOWS_PER_CTA": "16", "HB_GATE_WARPS": "16"},
    "per_chunk_tile8": {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "8"},
    "per_chunk_tile16_pingpong": {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "16"},
    "per_chunk_tile16_single": {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "16", "HB_NO_PINGPONG": "1"},
    "per_chunk_tile32_pingpong": {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "32"},
    "per_chunk_tile32_single": {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "32", "HB_NO_PINGPONG": "1"},
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_repeated_launches_reproduce_the_fp32_engine(variant, monkeypatch):
    from helen_b200.predictor import WindowPredictor
    for k in ENV_KEYS:
        monkeypatch.delenv(k, raising=False)
    batch, seq, features = 45, 250, 10
    sd = random_state_dict(features, seed=5)
    gen = torch.Generator().manual_seed(77)
    images = torch.randint(0, 256, (batch, seq, features), dtype=torch.uint8, generator=gen).cuda()
    ref_pred = WindowPredictor(sd, device=0)
    ref_pred.set_engine("fp32")
    ref = ref_pred.predict(images, return_probs=True)
    ref_pred.close()
    for k, v in VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    pred = WindowPredictor(sd, device=0)
    pred.set_engine("tensor")
    first = pred.predict(images, return_probs=True)
    bad = []
    for rep in range(REPS):
        got = pred.predict(images, return_probs=True)
        err = max(float((got[2] - ref[2]).abs().max()), float((got[3] - ref[3]).abs().max()))
        if err > 5e-6 or not (torch.equal(got[0], first[0]) and torch.equal(got[1], first[1])):
            bad.append((rep, err))
    pred.close()
    assert not bad, f"{variant}: {len(bad)} of {REPS} launches differ, first {bad[:4]}"


def test_repeated_full_size_batches_are_identical():
    """BASELINE configs[1] size: 40 launches of the same 256-window batch must give the same labels every time."""
    from helen_b200.predictor import WindowPredictor
    sd = random_state_dict(10, seed=0)
    gen = torch.Generator().manual_seed(1)
    images = torch.randint(0, 256, (256, 1000, 10), dtype=torch.uint8, generator=gen).cuda()
    pred = WindowPredictor(sd, device=0)
    base, rle = pred.predict(images)
    for rep in range(40):
        b, r = pred.predict(images)
        assert torch.equal(b, base) and torch.equal(r, rle), f"launch {rep} differs"
    pred.close()
