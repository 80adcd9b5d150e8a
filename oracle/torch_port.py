"""torch-CPU port of the reference's own CPU predict path (nn.GRU / nn.Linear on ATen).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  This is the CPU arm that
``bench.py`` times (``cpu_baseline`` and ``--impl reference``): the reference itself is a
Python package that cannot travel to the GPU box, so the arm is a *port* with the same
library calls the reference makes:

* module structure / parameter names -- ``helen/modules/python/models/TransducerModel.py:24-58``
  (two one-layer bidirectional batch_first ``nn.GRU`` + ``Linear(2H,5)`` + ``Linear(2H,11)``)
* forward -- ``TransducerModel.py:60-79``
* driver loop -- ``helen/modules/python/models/predict.py:90-154``

``tests/test_oracle.py`` checks it against fixtures minted from the real reference class.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from .explicit import HIDDEN, N_BASE, N_RLE, TRAIN_WINDOW, WINDOW_JUMP, chunk_starts, state_dict_shapes


class TransducerPort(nn.Module):
    """Same parameters, same state_dict keys, same forward as the reference model."""

    def __init__(self, image_features: int, hidden_size: int = HIDDEN,
                 num_base_classes: int = N_BASE, num_rle_classes: int = N_RLE):
        super().__init__()
        self.hidden_size = hidden_size
        self.gru_encoder = nn.GRU(image_features, hidden_size, num_layers=1,
                                  bidirectional=True, batch_first=True)
        self.gru_decoder = nn.GRU(2 * hidden_size, hidden_size, num_layers=1,
                                  bidirectional=True, batch_first=True)
        self.dense1_base = nn.Linear(2 * hidden_size, num_base_classes)
        self.dense2_rle = nn.Linear(2 * hidden_size, num_rle_classes)

    def forward(self, x, hidden):
        h0 = hidden.transpose(0, 1).contiguous()
        y1, h1 = self.gru_encoder(x, h0)
        y2, h2 = self.gru_decoder(y1, h1)
        return self.dense1_base(y2), self.dense2_rle(y2), h2.transpose(0, 1).contiguous()


def random_state_dict(image_features: int, seed: int, scale: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Random-init parameter set, uniform(-1/sqrt(H), 1/sqrt(H)) * scale, from a private
    generator (does not depend on torch's module-construction RNG order)."""
    gen = torch.Generator().manual_seed(seed)
    bound = scale / (HIDDEN ** 0.5)
    out = OrderedDict()
    for key, shape in state_dict_shapes(image_features).items():
        out[key] = (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound
    return out


@torch.no_grad()
def predict_port(model: nn.Module, images_u8: torch.Tensor,
                 window: int = TRAIN_WINDOW, jump: int = WINDOW_JUMP):
    """predict.py:90-154 on CPU tensors; images_u8 [B, T, F] uint8."""
    images = images_u8.type(torch.FloatTensor)
    batch, seq_len = images.size(0), images.size(1)
    param = next(model.parameters())
    images = images.to(param.dtype)
    hidden = torch.zeros(batch, 2, HIDDEN, dtype=param.dtype)
    p_base = torch.zeros(batch, seq_len, N_BASE, dtype=param.dtype)
    p_rle = torch.zeros(batch, seq_len, N_RLE, dtype=param.dtype)
    for i in chunk_starts(seq_len, window, jump):
        base, rle, hidden = model(images[:, i:i + window], hidden)
        pad = nn.ZeroPad2d((0, 0, i, seq_len - i - window))
        p_base = torch.add(p_base, pad(torch.softmax(base, dim=2)))
        p_rle = torch.add(p_rle, pad(torch.softmax(rle, dim=2)))
    _, base_label = torch.max(p_base, 2)
    _, rle_label = torch.max(p_rle, 2)
    return {
        "base_prob": p_base.numpy(),
        "rle_prob": p_rle.numpy(),
        "base_label": base_label.numpy().astype(np.uint8),
        "rle_label": rle_label.numpy().astype(np.uint8),
        "hidden": hidden.numpy(),
    }
