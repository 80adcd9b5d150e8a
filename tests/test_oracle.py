"""Oracle vs the golden fixtures minted from the real reference class (CPU only).

Tolerances: the two oracle implementations and the reference differ only in fp32
summation order, so accumulated probabilities must agree to 2e-5 abs and labels must be
identical wherever the reference's own top-1/top-2 margin is >= 1e-5.
"""
import numpy as np
import pytest
import torch

from conftest import golden_case_names, load_case
from oracle import OracleWeights, TransducerPort, forward_chunk, predict_port, predict_windows, chunk_starts
from oracle.explicit import top2_margin

MARGIN = 1e-5
PROB_TOL = 2e-5


def assert_labels_match(ref_prob, ref_label, got_label, what):
    diff = ref_label != got_label
    if diff.any():
        margins = top2_margin(ref_prob)[diff]
        assert (margins < MARGIN).all(), f"{what}: {int(diff.sum())} flips, worst margin {margins.max():.3e}"


@pytest.mark.parametrize("name", golden_case_names())
def test_explicit_oracle_matches_reference(name):
    case, state = load_case(name)
    if case["images"].shape[0] > 8:            # keep the CPU suite short: subsample big batches
        sel = slice(0, 8)
    else:
        sel = slice(None)
    out = predict_windows(OracleWeights.from_state_dict(state, np.float32), case["images"][sel])
    for head in ("base", "rle"):
        ref_prob = case[f"f32_{head}_prob"][sel]
        assert np.abs(out[f"{head}_prob"] - ref_prob).max() <= PROB_TOL
        assert_labels_match(ref_prob, case[f"f32_{head}_label"][sel], out[f"{head}_label"], f"{name}/{head}")
    assert np.abs(out["hidden"] - case["f32_hidden"][sel]).max() <= 1e-4


@pytest.mark.parametrize("name", ["F90_B2_T150_uniform", "F10_B3_T1000_pileup"])
def test_explicit_oracle_fp64_matches_reference_fp64(name):
    case, state = load_case(name)
    out = predict_windows(OracleWeights.from_state_dict(state, np.float64), case["images"])
    for head in ("base", "rle"):
        assert np.abs(out[f"{head}_prob"] - case[f"f64_{head}_prob"]).max() <= 1e-9
        assert (out[f"{head}_label"] == case[f"f64_{head}_label"]).all()


@pytest.mark.parametrize("name", golden_case_names())
def test_torch_port_matches_reference(name):
    case, state = load_case(name)
    model = TransducerPort(case["images"].shape[2]).eval()
    model.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    torch.set_num_threads(1)
    out = predict_port(model, torch.from_numpy(case["images"]))
    for head in ("base", "rle"):
        # same library calls as the reference on the same torch build: expect bit identity
        assert np.array_equal(out[f"{head}_prob"], case[f"f32_{head}_prob"])
        assert np.array_equal(out[f"{head}_label"], case[f"f32_{head}_label"])


def test_forward_chunk_matches_reference_first_chunk():
    case, state = load_case("cfg1_F10_B64_T100")
    w = OracleWeights.from_state_dict(state)
    x = case["images"][:8].astype(np.float32)
    base, rle, hidden = forward_chunk(w, x, np.zeros((8, 2, 128), np.float32))
    assert np.abs(base - case["f32_chunk0_base"][:8]).max() <= 1e-5
    assert np.abs(rle - case["f32_chunk0_rle"][:8]).max() <= 1e-5
    assert np.abs(hidden - case["f32_chunk0_hidden"][:8]).max() <= 1e-5


def test_module_prefix_is_stripped():
    case, state = load_case("F90_B1_T100_pileup")
    prefixed = {"module." + k: v for k, v in state.items()}
    a = predict_windows(OracleWeights.from_state_dict(state), case["images"])
    b = predict_windows(OracleWeights.from_state_dict(prefixed), case["images"])
    assert np.array_equal(a["base_prob"], b["base_prob"])


def test_chunk_starts():
    assert chunk_starts(1000) == list(range(0, 901, 50)) and len(chunk_starts(1000)) == 19
    assert chunk_starts(100) == [0]
    assert chunk_starts(150) == [0, 50]
    assert chunk_starts(99) == []
    assert chunk_starts(149) == [0]


def test_first_index_tie_break():
    # all-zero heads => every class ties; torch.max / the oracle must return index 0
    case, state = load_case("F90_B1_T100_pileup")
    state = {k: (np.zeros_like(v) if k.startswith("dense") else v) for k, v in state.items()}
    out = predict_windows(OracleWeights.from_state_dict(state), case["images"])
    assert (out["base_label"] == 0).all() and (out["rle_label"] == 0).all()
    assert np.allclose(out["base_prob"], 0.2) and np.allclose(out["rle_prob"], 1 / 11)


def test_port_autograd_reproduces_reference_training_fixture():
    """The training-step oracle (autograd of the torch port with the reference's two criteria, train.py:121-126)
    against the fixture minted from the reference model class (tests/golden/make_golden_train.py)."""
    import os
    from conftest import GOLDEN_DIR
    from oracle import TransducerPort
    fx = np.load(os.path.join(GOLDEN_DIR, "train_F10_B6_W100.npz"))
    state = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN_DIR, f"model_{str(fx['model'])}.npz")).items()}
    port = TransducerPort(10)
    port.load_state_dict(state)
    ob, orl, oh = port(torch.from_numpy(fx["x"]), torch.from_numpy(fx["hidden"]))
    lb, lr = torch.from_numpy(fx["label_base"]), torch.from_numpy(fx["label_rle"])
    loss_b = torch.nn.CrossEntropyLoss()(ob.reshape(-1, 5), lb.reshape(-1))
    loss_r = torch.nn.CrossEntropyLoss(weight=torch.from_numpy(fx["class_weights"]))(orl.reshape(-1, 11), lr.reshape(-1))
    (loss_b + loss_r).backward()
    np.testing.assert_allclose([loss_b.item() + loss_r.item(), loss_b.item(), loss_r.item()], fx["loss_f32"], rtol=1e-6)
    assert np.abs(oh.detach().numpy() - fx["hidden_out_f32"]).max() <= 1e-6
    for name, p in port.named_parameters():
        ref = fx[f"grad_f32/{name}"]
        got = p.grad.double().flatten()
        assert np.abs(got[torch.from_numpy(fx[f"grad_idx/{name}"])].numpy() - ref[2:]).max() <= 1e-6 * (np.abs(ref[2:]).max() + 1e-12), name
        # fp32 vs fp64 of the reference itself: the noise floor the CUDA path is compared under
        assert abs(ref[0] - fx[f"grad_f64/{name}"][0]) <= 1e-4 * fx[f"grad_f64/{name}"][0], name
