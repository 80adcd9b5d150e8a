"""Training step of one chunk (BASELINE configs[3]) against the fixture minted from the
reference model's autograd and against autograd of the oracle's torch port.

Tolerances: losses 1e-5 relative; every parameter gradient within 1e-4 of its own largest entry (fp32 CUDA vs
fp32 torch-CPU autograd differ by summation order; the fixture's fp32-vs-fp64 difference is reported)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import TransducerPort, random_state_dict

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-4
LOSS_TOL = 1e-5


def make_model(state, features):
    from helen_b200.models.TransducerModel import TransducerGRU
    model = TransducerGRU(1, features, 1, 128, 5, 11)
    model.load_state_dict(state)
    return model.cuda()


def test_train_step_matches_reference_fixture():
    from helen_b200.models.train_step import ChunkTrainer
    fx = np.load(os.path.join(GOLDEN_DIR, "train_F10_B6_W100.npz"))
    state = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN_DIR, f"model_{str(fx['model'])}.npz")).items()}
    model = make_model(state, 10)
    trainer = ChunkTrainer(model, class_weights=fx["class_weights"].tolist())
    loss, lb, lr, hidden = trainer.step(torch.from_numpy(fx["x"]).cuda(), torch.from_numpy(fx["hidden"]).cuda(),
                                        torch.from_numpy(fx["label_base"]).cuda(), torch.from_numpy(fx["label_rle"]).cuda())
    np.testing.assert_allclose([loss, lb, lr], fx["loss_f32"], rtol=LOSS_TOL)
    assert np.abs(hidden.cpu().numpy() - fx["hidden_out_f32"]).max() <= 2e-5
    for name, p in model.named_parameters():
        ref = fx[f"grad_f32/{name}"]
        got = p.grad.detach().double().cpu().flatten()
        scale = np.abs(ref[2:]).max() + 1e-12
        idx = torch.from_numpy(fx[f"grad_idx/{name}"])
        assert np.abs(got[idx].numpy() - ref[2:]).max() <= GRAD_TOL * max(scale, abs(ref[0]) / np.sqrt(got.numel())), name
        assert abs(got.norm().item() - ref[0]) <= GRAD_TOL * ref[0] + 1e-9, name
    trainer.close()


@pytest.mark.parametrize("batch,width,features,with_hidden", [(5, 100, 10, True), (3, 37, 90, False), (9, 64, 33, True)])
def test_train_step_matches_port_autograd(batch, width, features, with_hidden):
    from helen_b200.models.train_step import ChunkTrainer
    from helen_b200.options import TrainOptions
    sd = random_state_dict(features, seed=features + batch)
    gen = torch.Generator().manual_seed(width)
    x = torch.randint(0, 256, (batch, width, features), generator=gen).float()
    hidden = (torch.rand(batch, 2, 128, generator=gen) - 0.5) if with_hidden else torch.zeros(batch, 2, 128)
    lb = torch.randint(0, 5, (batch, width), generator=gen)
    lr = torch.randint(0, 11, (batch, width), generator=gen)
    # oracle: the port's autograd with the reference's two criteria (train.py:121-126)
    port = TransducerPort(features)
    port.load_state_dict(sd)
    ob, orl, oh = port(x, hidden)
    loss_b = torch.nn.CrossEntropyLoss()(ob.reshape(-1, 5), lb.reshape(-1))
    loss_r = torch.nn.CrossEntropyLoss(weight=torch.tensor(TrainOptions.CLASS_WEIGHTS))(orl.reshape(-1, 11), lr.reshape(-1))
    (loss_b + loss_r).backward()
    model = make_model(sd, features)
    trainer = ChunkTrainer(model)
    loss, l_b, l_r, h_out, base, rle = trainer.step(x.cuda(), hidden.cuda() if with_hidden else None, lb.cuda(), lr.cuda(), return_logits=True)
    np.testing.assert_allclose([l_b, l_r], [loss_b.item(), loss_r.item()], rtol=LOSS_TOL)
    assert abs(loss - (loss_b.item() + loss_r.item())) <= LOSS_TOL * abs(loss)
    assert (base.cpu() - ob.detach()).abs().max() <= 2e-5 and (rle.cpu() - orl.detach()).abs().max() <= 2e-5
    assert (h_out.cpu() - oh.detach()).abs().max() <= 2e-5
    for (name, p), (pname, q) in zip(model.named_parameters(), port.named_parameters()):
        assert name == pname
        err = (p.grad.cpu() - q.grad).abs().max().item()
        assert err <= GRAD_TOL * q.grad.abs().max().item() + 1e-9, f"{name}: {err:.3e} vs max {q.grad.abs().max().item():.3e}"
    # a second step overwrites (not accumulates) the gradients and an optimizer can consume them
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    before = [p.detach().clone() for p in model.parameters()]
    loss2, *_ = trainer.step(x.cuda(), hidden.cuda() if with_hidden else None, lb.cuda(), lr.cuda())
    assert abs(loss2 - loss) <= 1e-6 * abs(loss)
    opt.step()
    assert all(not torch.equal(b, p.detach()) for b, p in zip(before, model.parameters()))
    loss3, *_ = trainer.step(x.cuda(), hidden.cuda() if with_hidden else None, lb.cuda(), lr.cuda())
    assert loss3 < loss
    trainer.close()


def test_label_out_of_range_is_a_value_error():
    """nn.CrossEntropyLoss raises a device assert for such a label (train.py:121-126); here the step answers ValueError
    and nothing is read outside the class-weight table."""
    from helen_b200.models.train_step import ChunkTrainer
    model = make_model(random_state_dict(10, seed=3), 10)
    trainer = ChunkTrainer(model)
    x = torch.randint(0, 256, (2, 20, 10)).float().cuda()
    lb = torch.zeros(2, 20, dtype=torch.int64).cuda()
    for bad_base, bad_rle in ((5, 0), (0, 11), (-1, 0), (0, 1 << 40)):
        b, r = lb.clone(), lb.clone()
        b[1, 7], r[0, 3] = bad_base, bad_rle
        with pytest.raises(ValueError, match="labels out of range"):
            trainer.step(x, None, b, r)
    loss, *_ = trainer.step(x, None, lb, lb)
    assert np.isfinite(loss)
    trainer.close()


def test_training_loop_mirror(tmp_path, monkeypatch):
    """helen_b200.models.train.train (train.py:19-259 mirror) on a tiny in-memory data set: the training loss falls from
    epoch to epoch, every epoch leaves a reference-format checkpoint that load_simple_model reads back, and retraining
    continues from it."""
    import fake_h5
    from helen_b200 import hdf5
    from helen_b200.models.ModelHander import ModelHandler
    from helen_b200.models.train import train
    fake_h5.reset()
    monkeypatch.setattr(hdf5, "open_file", fake_h5.open_file)
    data_dir = tmp_path / "images"
    data_dir.mkdir()
    (data_dir / "train.h5").write_bytes(b"")                  # the directory listing is real, the contents are the fake store
    path = str(data_dir / "train.h5")
    rng = np.random.default_rng(3)
    for i in range(6):
        img = rng.integers(0, 256, (200, 10), dtype=np.uint8)
        fake_h5.add_image(path, f"img{i}", "chr1", 0, 200, i, img, np.zeros((200, 3), int))
        # learnable labels: a function of the pixels
        fake_h5.add_labels(path, f"img{i}", (img[:, 0] // 52).clip(0, 4), (img[:, 1] // 24).clip(0, 10))
    model_dir, stats_dir = str(tmp_path) + "/models_", str(tmp_path) + "/stats_"
    torch.manual_seed(0)
    model, optimizer, stats = train(str(data_dir), str(data_dir), batch_size=3, epoch_limit=3, gpu_mode=True, num_workers=0,
                                    retrain_model=False, retrain_model_path=None, gru_layers=1, hidden_size=128, lr=1e-3, decay=0.0,
                                    model_dir=model_dir, stats_dir=stats_dir, not_hyperband=True)
    losses = [l for _, l in stats['loss_epoch']]
    assert len(losses) == 3 and losses[-1] < losses[0], losses
    loaded, hidden, layers, epochs = ModelHandler.load_simple_model(model_dir + "HELEN_epoch_3_checkpoint.pkl", 1, 10, 1000, 5, 11)
    assert (hidden, layers, epochs) == (128, 1, 2)
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), loaded.state_dict()[k])
    _, _, stats2 = train(str(data_dir), str(data_dir), 3, 1, True, 0, True, model_dir + "HELEN_epoch_3_checkpoint.pkl", 1, 128, 1e-3, 0.0,
                         model_dir, stats_dir, True)
    assert stats2['loss_epoch'][-1][1] < losses[-1] * 1.05
    fake_h5.reset()
