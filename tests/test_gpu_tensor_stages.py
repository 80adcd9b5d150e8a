"""Stage-by-stage bring-up checks of the tcgen05 engine against the fp32 engine on the same
device (tighter than the oracle comparison: both run the same algorithm, so accumulated
probabilities must agree to 5e-6)."""
import numpy as np
import pytest
import torch

from oracle import random_state_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("engine", ["tensor"])
@pytest.mark.parametrize("batch,seq,features", [(3, 100, 10), (16, 150, 10), (21, 200, 90), (40, 100, 10), (5, 137, 10), (75, 250, 33)])
def test_stage_matches_fp32_engine(engine, batch, seq, features):
    from helen_b200.predictor import WindowPredictor
    sd = random_state_dict(features, seed=features + 1)
    gen = torch.Generator().manual_seed(batch)
    images = torch.randint(0, 256, (batch, seq, features), dtype=torch.uint8, generator=gen).cuda()
    pred = WindowPredictor(sd, device=0)
    pred.set_engine("fp32")
    ref = [t.cpu().numpy() for t in pred.predict(images, return_probs=True)]
    pred.set_engine(engine)
    got = [t.cpu().numpy() for t in pred.predict(images, return_probs=True)]
    pred.close()
    err = max(np.abs(got[2] - ref[2]).max(), np.abs(got[3] - ref[3]).max())
    assert np.isfinite(got[2]).all() and np.isfinite(got[3]).all()
    assert err <= 5e-6, f"{engine}: max |dP| vs fp32 engine = {err:.3e}"


VARIANTS = {
    "tile8_stacked": {"HB_WINDOWS_PER_CTA": "8"},
    "tile8_3term": {"HB_WINDOWS_PER_CTA": "8", "HB_NO_STACK": "1"},
    "tile16_stacked": {"HB_WINDOWS_PER_CTA": "16"},
    "tile16_3term": {"HB_WINDOWS_PER_CTA": "16", "HB_NO_STACK": "1"},
    "tile32": {"HB_WINDOWS_PER_CTA": "32"},
    "no_pair_no_pdl": {"HB_NO_PAIR": "1", "HB_NO_PDL": "1"},
    "per_chunk_launches": {"HB_NO_CHUNKLOOP": "1"},
    "per_chunk_tile8_3term": {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "8", "HB_NO_STACK": "1"},
    "per_chunk_tile16": {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "16"},
    "per_chunk_tile32": {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "32"},
    "per_chunk_tile16_single_tile": {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "16", "HB_NO_PINGPONG": "1"},
    "per_chunk_tile32_single_tile": {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "32", "HB_NO_PINGPONG": "1"},
    "chunkloop_no_pair": {"HB_NO_PAIR": "1"},
    "chunkloop_no_pixel_jobs": {"HB_NO_PIXEL_JOBS": "1"},
    "chunkloop_tile16_pixel_jobs": {"HB_WINDOWS_PER_CTA": "16"},
    "chunkloop_tile16_one_tile": {"HB_WINDOWS_PER_CTA": "16", "HB_NO_LOOP_PINGPONG": "1"},
    "chunkloop_pixel_jobs_first": {"HB_PIXELS_FIRST": "1"},
    "chunkloop_tile16_one_tile_8_gate_warps": {"HB_WINDOWS_PER_CTA": "16", "HB_NO_LOOP_PINGPONG": "1", "HB_GATE_WARPS": "8"},
    "chunkloop_few_heads_workers": {"HB_HEADS_WORKERS": "2"},
    "chunkloop_16_gate_warps": {"HB_GATE_WARPS": "16"},
    "chunkloop_16_gate_warps_tile16": {"HB_GATE_WARPS": "16", "HB_WINDOWS_PER_CTA": "16"},
    "chunkloop_not_cooperative": {"HB_NO_COOPERATIVE": "1"},
    "chunkloop_many_heads_workers": {"HB_HEADS_WORKERS": "24"},
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_kernel_variants_match_fp32_engine(variant, monkeypatch):
    """Every recurrence tile / launch-structure variant of the tensor engine (the switches are read from the
    environment when the handle is created) against the fp32 engine, same tolerance as above."""
    from helen_b200.predictor import WindowPredictor
    for k in ("HB_WINDOWS_PER_CTA", "HB_NO_STACK", "HB_NO_PAIR", "HB_NO_PDL", "HB_NO_CHUNKLOOP", "HB_HEADS_WORKERS", "HB_NO_LIVE8", "HB_NO_PIXEL_JOBS", "HB_NO_PINGPONG",
              "HB_GATE_WARPS", "HB_NO_COOPERATIVE", "HB_NO_LOOP_PINGPONG", "HB_PIXELS_FIRST"):
        monkeypatch.delenv(k, raising=False)
    batch, seq, features = 45, 250, 10
    sd = random_state_dict(features, seed=5)
    gen = torch.Generator().manual_seed(77)
    images = torch.randint(0, 256, (batch, seq, features), dtype=torch.uint8, generator=gen).cuda()
    ref_pred = WindowPredictor(sd, device=0)
    ref_pred.set_engine("fp32")
    ref = [t.cpu().numpy() for t in ref_pred.predict(images, return_probs=True)]
    ref_pred.close()
    for k, v in VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    pred = WindowPredictor(sd, device=0)
    pred.set_engine("tensor")
    got = [t.cpu().numpy() for t in pred.predict(images, return_probs=True)]
    again = [t.cpu().numpy() for t in pred.predict(images, return_probs=True)]
    pred.close()
    err = max(np.abs(got[2] - ref[2]).max(), np.abs(got[3] - ref[3]).max())
    assert np.isfinite(got[2]).all() and np.isfinite(got[3]).all()
    assert err <= 5e-6, f"{variant}: max |dP| vs fp32 engine = {err:.3e}"
    assert np.array_equal(got[0], again[0]) and np.array_equal(got[1], again[1]), f"{variant}: not repeatable"
