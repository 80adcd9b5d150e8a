#!/usr/bin/env python
"""Mint the training-step fixture from the REAL reference model class (build container only):

    python tests/golden/make_golden_train.py

One chunk step of helen/modules/python/models/train.py:189-201 -- forward of the reference ``TransducerGRU``,
``CrossEntropyLoss()`` on the base head plus ``CrossEntropyLoss(weight=CLASS_WEIGHTS)`` on the run-length head
(train.py:121-126, Options.py:29), ``loss.backward()`` -- on seeded inputs with a non-zero initial hidden state and the parameters of ``model_F10_seed0.npz``.
Stores the inputs, the three loss values, the returned hidden state, and for every parameter gradient its L2 norm,
its sum and 512 seeded sample entries (the full gradients would be 1.6 MB), computed in fp32 and in fp64.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REFERENCE = os.environ.get("HELEN_REFERENCE", "/root/reference")
sys.path.insert(0, REFERENCE)
from helen.modules.python.models.TransducerModel import TransducerGRU  # noqa: E402
from helen.modules.python.Options import TrainOptions  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SAMPLES = 512


def one_step(model, x, hidden, label_base, label_rle, dtype):
    model = model.to(dtype)
    for p in model.parameters():
        p.grad = None
    criterion_base = nn.CrossEntropyLoss()
    criterion_rle = nn.CrossEntropyLoss(weight=torch.Tensor(TrainOptions.CLASS_WEIGHTS).to(dtype))
    output_base, output_rle, hidden_out = model(x.to(dtype), hidden.to(dtype))
    loss_base = criterion_base(output_base.contiguous().view(-1, 5), label_base.contiguous().view(-1))
    loss_rle = criterion_rle(output_rle.contiguous().view(-1, 11), label_rle.contiguous().view(-1))
    loss = loss_base + loss_rle
    loss.backward()
    return loss.item(), loss_base.item(), loss_rle.item(), hidden_out.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def main():
    features, batch, width = 10, 6, 100
    model = TransducerGRU(1, features, 1, 128, 5, 11)
    state = np.load(os.path.join(HERE, "model_F10_seed0.npz"))       # the committed parameter fixture (make_golden.py)
    model.load_state_dict({k: torch.from_numpy(state[k]) for k in state.files})
    gen = torch.Generator().manual_seed(5)
    x = torch.randint(0, 256, (batch, width, features), generator=gen).float()
    hidden = (torch.rand(batch, 2, 128, generator=gen) * 2 - 1) * 0.5
    label_base = torch.randint(0, 5, (batch, width), generator=gen)
    label_rle = torch.randint(0, 11, (batch, width), generator=gen)
    out = {"x": x.numpy(), "hidden": hidden.numpy(), "label_base": label_base.numpy(), "label_rle": label_rle.numpy(),
           "class_weights": np.asarray(TrainOptions.CLASS_WEIGHTS, np.float32)}
    out["model"] = np.asarray("F10_seed0")
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        loss, lb, lr, h, grads = one_step(model, x, hidden, label_base, label_rle, dtype)
        out[f"loss_{tag}"] = np.asarray([loss, lb, lr], np.float64)
        out[f"hidden_out_{tag}"] = h.double().numpy()
        for k, g in grads.items():
            flat = g.double().flatten()
            idx = torch.randperm(flat.numel(), generator=torch.Generator().manual_seed(len(k)))[:SAMPLES]
            out[f"grad_idx/{k}"] = idx.numpy()
            out[f"grad_{tag}/{k}"] = np.concatenate([[flat.norm().item(), flat.sum().item()], flat[idx].numpy()])
    path = os.path.join(HERE, "train_F10_B6_W100.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; loss", out["loss_f32"])


if __name__ == "__main__":
    main()
