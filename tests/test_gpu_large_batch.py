"""GPU parity at the sizes that pick each launch structure of the tensor engine by themselves.

`helen polish` defaults to --batch_size 512 (reference helen/helen.py:32-39).  B <= 320 runs the chunk-loop kernel with
8-window recurrence tiles, 320 < B <= 512 the chunk-loop kernel with two 8-window tiles per recurrence CTA
(tc_chunkloop2_kernel), larger batches per-chunk launches (two-tile recurrence kernels).  Nothing here forces a variant: the
library chooses, the test records what it chose (hb_last_launch_plan) and compares with the CPU oracle
(oracle.predict_port == the reference's own nn.GRU calls, predict.py:90-154) on

  * ALL windows of BASELINE configs[1] (B=256, T=1000) at F=10 and F=90, and
  * >= 40 sampled windows of every larger batch: the first tile, the last (partial) tile, and an even spread that
    touches every recurrence CTA wave.

Gate (same as tests/test_gpu_parity.py): accumulated probabilities within 2e-5 abs, labels identical wherever the
oracle's top-1/top-2 margin is >= 1e-5.  Every case reports `checked / flips_sub_margin / flips_above_margin`
(SURVEY 8d "count + list") through conftest.parity_report.
"""
import numpy as np
import pytest
import torch

from conftest import parity_report
from oracle import TransducerPort, predict_port, random_state_dict
from oracle.explicit import top2_margin

pytestmark = pytest.mark.gpu

PROB_TOL = 2e-5
MARGIN = 1e-5
T = 1000


def oracle_for(sd, features, images_u8):
    model = TransducerPort(features).eval()
    model.load_state_dict(sd)
    return predict_port(model, images_u8)


def flip_census(ref, labels_b, labels_r, prob_b=None, prob_r=None):
    """Counts of label differences against the oracle, split by the oracle's own top-1/top-2 margin."""
    out = {"positions": int(labels_b.size + labels_r.size), "flips_sub_margin": 0, "flips_above_margin": 0, "flip_list": []}
    for head, got, ref_lab, ref_prob in (("base", labels_b, ref["base_label"], ref["base_prob"]),
                                         ("rle", labels_r, ref["rle_label"], ref["rle_prob"])):
        diff = np.argwhere(got != ref_lab)
        if len(diff):
            margins = top2_margin(ref_prob)
            for w, t in diff:
                m = float(margins[w, t])
                out["flips_sub_margin" if m < MARGIN else "flips_above_margin"] += 1
                if len(out["flip_list"]) < 16:
                    out["flip_list"].append({"head": head, "window": int(w), "column": int(t), "oracle_margin": m})
    if prob_b is not None:
        out["max_abs_dP"] = float(max(np.abs(prob_b - ref["base_prob"]).max(), np.abs(prob_r - ref["rle_prob"]).max()))
    return out


def sample_windows(batch, tile):
    """First tile, last tile (partial when batch % tile != 0) and an even spread over every CTA wave."""
    first = list(range(min(tile, batch)))
    last = list(range(max(0, (batch - 1) // tile * tile), batch))
    spread = list(np.linspace(0, batch - 1, 32).astype(int))
    return sorted(set(first[:8] + last[-8:] + spread))


@pytest.mark.parametrize("features", [10, 90])
def test_config2_every_window_matches_oracle(features):
    """BASELINE configs[1] literally: B=256, T=1000, model seed 0, images seed 1 -- all 256 windows against the oracle."""
    from helen_b200.predictor import WindowPredictor
    sd = random_state_dict(features, seed=0)
    gen = torch.Generator().manual_seed(1)
    images = torch.randint(0, 256, (256, T, features), dtype=torch.uint8, generator=gen)
    ref = oracle_for(sd, features, images)
    pred = WindowPredictor(sd, device=0)
    dev = images.cuda()
    base, rle = pred.predict(dev)                                    # the product path (labels only)
    plan = pred.last_launch_plan()
    base_p, rle_p, pb, pr = pred.predict(dev, return_probs=True)     # the same batch with the probabilities returned
    torch.cuda.synchronize()
    assert torch.equal(base, base_p) and torch.equal(rle, rle_p)
    census = flip_census(ref, base.cpu().numpy(), rle.cpu().numpy(), pb.cpu().numpy(), pr.cpu().numpy())
    parity_report({"case": f"config2_B256_T{T}_F{features}", "windows_checked": 256, "plan": plan, **census,
                   "oracle": f"torch {torch.__version__} CPU fp32 port"})
    pred.close()
    assert plan["chunkloop"] == 1, plan
    assert census["max_abs_dP"] <= PROB_TOL, census
    assert census["flips_above_margin"] == 0, census


@pytest.mark.parametrize("batch,features", [(384, 10), (500, 10), (512, 10), (1024, 10), (2048, 10),
                                            (384, 90), (512, 90), (1000, 90), (2048, 90)])
def test_auto_selected_large_batch_matches_oracle(batch, features):
    """Batches above the 8-window chunk-loop kernel's reach: the two-tile chunk-loop kernel up to 512 windows, above that
    per-chunk launches with the two-tile recurrence kernels, multi-wave projection grids and gi images of up to 6.3 GB:
    checked here at their own sizes, against the oracle."""
    from helen_b200.predictor import WindowPredictor
    sd = random_state_dict(features, seed=batch % 7)
    gen = torch.Generator().manual_seed(batch + features)
    images = torch.randint(0, 256, (batch, T, features), dtype=torch.uint8, generator=gen)
    pred = WindowPredictor(sd, device=0)
    dev = images.cuda()
    base, rle = pred.predict(dev)
    plan = pred.last_launch_plan()
    base_p, rle_p, pb, pr = pred.predict(dev, return_probs=True)
    base2, rle2 = pred.predict(dev)                                  # run to run
    torch.cuda.synchronize()
    assert plan["chunkloop"] == (1 if batch <= 512 else 0), plan
    assert batch > 512 or plan["windows_per_cta"] == 16, plan
    assert torch.equal(base, base_p) and torch.equal(rle, rle_p)
    assert torch.equal(base, base2) and torch.equal(rle, rle2)
    idx = sample_windows(batch, max(plan["windows_per_cta"], 8))
    assert len(idx) >= 40
    ref = oracle_for(sd, features, images[idx])
    sel = torch.tensor(idx, device="cuda")
    census = flip_census(ref, base[sel].cpu().numpy(), rle[sel].cpu().numpy(), pb[sel].cpu().numpy(), pr[sel].cpu().numpy())
    parity_report({"case": f"auto_B{batch}_T{T}_F{features}", "windows_checked": len(idx), "plan": plan, **census,
                   "oracle": f"torch {torch.__version__} CPU fp32 port"})
    # windows are independent: the sampled windows predicted on their own (a small batch, chunk-loop kernel) must give
    # the labels they got inside the large batch
    small_b, small_r = pred.predict(dev[sel].contiguous())
    assert torch.equal(small_b, base[sel]) and torch.equal(small_r, rle[sel])
    pred.close()
    assert census["max_abs_dP"] <= PROB_TOL, census
    assert census["flips_above_margin"] == 0, census


def test_two_handles_on_one_device_predict_concurrently():
    """Two handles on the same GPU, each on its own stream, with launches interleaved: neither may trap or disturb the
    other (the chunk-loop kernel's CTAs wait for each other through global counters, so its launch has to be
    co-resident as a whole)."""
    from helen_b200.predictor import WindowPredictor
    sd_a, sd_b = random_state_dict(10, seed=11), random_state_dict(10, seed=12)
    gen = torch.Generator().manual_seed(5)
    img_a = torch.randint(0, 256, (96, 400, 10), dtype=torch.uint8, generator=gen).cuda()
    img_b = torch.randint(0, 256, (160, 400, 10), dtype=torch.uint8, generator=gen).cuda()
    pa, pb_ = WindowPredictor(sd_a, device=0), WindowPredictor(sd_b, device=0)
    want_a, want_b = pa.predict(img_a), pb_.predict(img_b)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    got_a, got_b = [], []
    for _ in range(6):
        with torch.cuda.stream(sa):
            got_a.append(pa.predict(img_a))
        with torch.cuda.stream(sb):
            got_b.append(pb_.predict(img_b))
    torch.cuda.synchronize()
    for g in got_a:
        assert torch.equal(g[0], want_a[0]) and torch.equal(g[1], want_a[1])
    for g in got_b:
        assert torch.equal(g[0], want_b[0]) and torch.equal(g[1], want_b[1])
    pa.close()
    pb_.close()
