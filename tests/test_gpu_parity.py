"""GPU parity tests: the CUDA path (through the C ABI) against the golden fixtures minted
from the reference model and against the CPU oracle.

Tolerances (fp32 path, north_star): accumulated probabilities within 2e-5 abs of the
reference fp32 result; chunk logits within 2e-5 abs; labels identical wherever the
reference's top-1/top-2 margin >= 1e-5 (sub-margin positions are counted and reported).
"""
import numpy as np
import pytest
import torch

from conftest import golden_case_names, load_case
from oracle import OracleWeights, predict_windows, random_state_dict, forward_chunk
from oracle.explicit import top2_margin

pytestmark = pytest.mark.gpu

PROB_TOL = 2e-5
LOGIT_TOL = 2e-5
MARGIN = 1e-5


def engines(pred):
    out = []
    for e in ("tensor", "fp32"):
        try:
            pred.set_engine(e)
            out.append(e)
        except ValueError:
            pass
    return out


@pytest.fixture(scope="module")
def predictor_cache():
    from helen_b200.predictor import WindowPredictor
    cache = {}

    def get(model_name, state):
        if model_name not in cache:
            cache[model_name] = WindowPredictor({k: torch.from_numpy(v) for k, v in state.items()}, device=0)
        return cache[model_name]

    yield get
    for p in cache.values():
        p.close()


def check_against(ref_prob_b, ref_prob_r, ref_lab_b, ref_lab_r, got, what):
    base, rle, pb, pr = [t.cpu().numpy() for t in got]
    assert np.isfinite(pb).all() and np.isfinite(pr).all()
    eb, er = np.abs(pb - ref_prob_b).max(), np.abs(pr - ref_prob_r).max()
    assert eb <= PROB_TOL and er <= PROB_TOL, f"{what}: prob error base {eb:.3e} rle {er:.3e}"
    for lab, ref_lab, ref_prob, head in ((base, ref_lab_b, ref_prob_b, "base"), (rle, ref_lab_r, ref_prob_r, "rle")):
        diff = lab != ref_lab
        if diff.any():
            m = top2_margin(ref_prob)[diff]
            assert (m < MARGIN).all(), f"{what}/{head}: {int(diff.sum())} label flips, worst margin {m.max():.3e}"
    # the labels must be the first-index argmax of the probabilities the kernel itself reports
    assert np.array_equal(base, pb.argmax(2).astype(np.uint8))
    assert np.array_equal(rle, pr.argmax(2).astype(np.uint8))


@pytest.mark.parametrize("name", golden_case_names())
def test_predict_matches_reference_golden(name, predictor_cache):
    case, state = load_case(name)
    pred = predictor_cache(str(case["model"]), state)
    images = torch.from_numpy(case["images"]).cuda()
    for engine in engines(pred):
        got = pred.predict(images, return_probs=True)
        torch.cuda.synchronize()
        check_against(case["f32_base_prob"], case["f32_rle_prob"], case["f32_base_label"], case["f32_rle_label"],
                      got, f"{name}[{engine}]")
        # labels-only call (no probability outputs) must give the same labels
        base2, rle2 = pred.predict(images)
        assert torch.equal(base2, got[0]) and torch.equal(rle2, got[1])


@pytest.mark.parametrize("name", ["cfg1_F10_B64_T100", "F90_B1_T100_pileup"])
def test_forward_chunk_matches_reference_logits(name, predictor_cache):
    case, state = load_case(name)
    pred = predictor_cache(str(case["model"]), state)
    x = torch.from_numpy(case["images"][:, :100].astype(np.float32)).cuda()
    hidden = torch.zeros(x.shape[0], 2, 128, device="cuda")
    base, rle, h = pred.forward_chunk(x, hidden)
    assert np.abs(base.cpu().numpy() - case["f32_chunk0_base"]).max() <= LOGIT_TOL
    assert np.abs(rle.cpu().numpy() - case["f32_chunk0_rle"]).max() <= LOGIT_TOL
    assert np.abs(h.cpu().numpy() - case["f32_chunk0_hidden"]).max() <= LOGIT_TOL


def test_forward_chunk_nonzero_hidden_and_odd_width():
    from helen_b200.predictor import WindowPredictor
    sd = random_state_dict(23, seed=5)
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, (5, 37, 23)).astype(np.float32)
    hidden = rng.uniform(-1, 1, (5, 2, 128)).astype(np.float32)
    rb, rr, rh = forward_chunk(OracleWeights.from_state_dict(sd), x, hidden)
    pred = WindowPredictor(sd, device=0)
    base, rle, h = pred.forward_chunk(torch.from_numpy(x).cuda(), torch.from_numpy(hidden).cuda())
    assert np.abs(base.cpu().numpy() - rb).max() <= LOGIT_TOL
    assert np.abs(rle.cpu().numpy() - rr).max() <= LOGIT_TOL
    assert np.abs(h.cpu().numpy() - rh).max() <= LOGIT_TOL
    pred.close()


def test_transducer_module_drop_in():
    """The nn.Module surface: load a state_dict, call model(x, hidden) like predict_gpu.py:129."""
    from helen_b200.models.TransducerModel import TransducerGRU
    case, state = load_case("cfg1_F10_B64_T100")
    model = TransducerGRU(1, 10, 1, 128, 5, 11)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    x = torch.from_numpy(case["images"][:16].astype(np.float32)).cuda()
    base, rle, h = model(x, torch.zeros(16, 2, 128).cuda())
    assert np.abs(base.cpu().numpy() - case["f32_chunk0_base"][:16]).max() <= LOGIT_TOL
    assert tuple(h.shape) == (16, 2, 128)


@pytest.mark.parametrize("batch,seq,features", [(1, 100, 10), (7, 149, 10), (5, 1000, 90), (2, 99, 10), (0, 1000, 10),
                                                (9, 450, 33), (130, 200, 10)])
def test_predict_matches_oracle_shapes(batch, seq, features):
    """Ragged / empty / odd sizes against the oracle on seeded inputs."""
    from helen_b200.predictor import WindowPredictor
    sd = random_state_dict(features, seed=features)
    gen = torch.Generator().manual_seed(batch * 1000 + seq)
    images = torch.randint(0, 256, (batch, seq, features), dtype=torch.uint8, generator=gen)
    pred = WindowPredictor(sd, device=0)
    for engine in engines(pred):
        got = pred.predict(images.cuda(), return_probs=True)
        torch.cuda.synchronize()
        if batch == 0:
            assert got[0].shape == (0, seq)
            continue
        ref = predict_windows(OracleWeights.from_state_dict(sd), images.numpy())
        check_against(ref["base_prob"], ref["rle_prob"], ref["base_label"], ref["rle_label"], got,
                      f"B{batch}_T{seq}_F{features}[{engine}]")
    pred.close()


def test_predict_host_entry_matches_device_entry():
    from helen_b200.predictor import WindowPredictor
    sd = random_state_dict(10, seed=3)
    gen = torch.Generator().manual_seed(11)
    images = torch.randint(0, 256, (12, 300, 10), dtype=torch.uint8, generator=gen)
    pred = WindowPredictor(sd, device=0)
    base_d, rle_d = pred.predict(images.cuda())
    base_h, rle_h, pb, pr = pred.predict_host(images.numpy(), return_probs=True)
    assert np.array_equal(base_h, base_d.cpu().numpy()) and np.array_equal(rle_h, rle_d.cpu().numpy())
    assert np.array_equal(base_h, pb.argmax(2)) and np.array_equal(rle_h, pr.argmax(2))
    pred.close()


def test_full_size_batch_properties():
    """BASELINE config 2 size (B=256, T=1000, F=10): size-independent properties.
    (a) windows are independent: any sub-batch gives identical labels; (b) a window's labels do
    not depend on its position in the batch; (c) the oracle agrees on a sample of windows."""
    from helen_b200.predictor import WindowPredictor
    sd = random_state_dict(10, seed=0)
    gen = torch.Generator().manual_seed(1)
    images = torch.randint(0, 256, (256, 1000, 10), dtype=torch.uint8, generator=gen)
    pred = WindowPredictor(sd, device=0)
    dev = images.cuda()
    for engine in engines(pred):
        base, rle, pb, pr = pred.predict(dev, return_probs=True)
        sub_b, sub_r = pred.predict(dev[37:101].contiguous())
        assert torch.equal(sub_b, base[37:101]) and torch.equal(sub_r, rle[37:101])
        perm = torch.randperm(256, generator=gen)
        pb2, pr2 = pred.predict(dev[perm.cuda()].contiguous())
        assert torch.equal(pb2, base[perm.cuda()]) and torch.equal(pr2, rle[perm.cuda()])
        # (d) run-to-run: the roles of the chunk-loop kernel hand work over through counters in global memory and pick
        # their job order at run time; results must not depend on that timing
        for _ in range(12):
            b_again, r_again, pb_again, pr_again = pred.predict(dev, return_probs=True)
            assert torch.equal(b_again, base) and torch.equal(r_again, rle)
            assert torch.equal(pb_again, pb) and torch.equal(pr_again, pr)
        sample = [0, 100, 255]
        ref = predict_windows(OracleWeights.from_state_dict(sd), images[sample].numpy())
        got = (base[sample], rle[sample], pb[sample], pr[sample])
        check_against(ref["base_prob"], ref["rle_prob"], ref["base_label"], ref["rle_label"], got, f"full[{engine}]")
    pred.close()


def test_errors_are_loud():
    from helen_b200.predictor import WindowPredictor
    sd = random_state_dict(10, seed=0)
    pred = WindowPredictor(sd, device=0)
    with pytest.raises(ValueError):
        pred.predict(torch.zeros(2, 100, 10, dtype=torch.uint8))            # CPU tensor
    with pytest.raises(ValueError):
        pred.predict(torch.zeros(2, 100, 11, dtype=torch.uint8).cuda())     # wrong F
    with pytest.raises(ValueError):
        pred.predict(torch.zeros(2, 100, 10, dtype=torch.float32).cuda())   # wrong dtype
    with pytest.raises(ValueError):
        pred.predict(torch.zeros(2, 100, 10, dtype=torch.uint8).cuda(), window=0)
    pred.close()
