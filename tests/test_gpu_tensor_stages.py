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
