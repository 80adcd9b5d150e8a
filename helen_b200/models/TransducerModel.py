"""TransducerGRU with the reference's constructor, parameter names and forward signature
(helen/modules/python/models/TransducerModel.py:20-93), computing through the CUDA library.

Parameters live in ordinary nn.Parameter tensors under the reference's state_dict keys
(``gru_encoder.weight_ih_l0`` ...), so reference ``.pkl`` checkpoints load unchanged.
``forward`` packs them into a native handle on first use for the device the inputs are on
and re-packs if the parameters change.  ``forward`` is the inference call (no autograd graph); the
training step over the same parameters is ``train_step.ChunkTrainer``.  There is no CPU execution path.
"""
import math

import torch
import torch.nn as nn

from ..predictor import WindowPredictor


class _GRUParameters(nn.Module):
    """Parameter container with torch.nn.GRU's names for a 1-layer bidirectional GRU."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        bound = 1.0 / math.sqrt(hidden_size)
        for rev in ("", "_reverse"):
            for name, shape in (("weight_ih_l0", (3 * hidden_size, input_size)),
                                ("weight_hh_l0", (3 * hidden_size, hidden_size)),
                                ("bias_ih_l0", (3 * hidden_size,)),
                                ("bias_hh_l0", (3 * hidden_size,))):
                p = nn.Parameter(torch.empty(shape).uniform_(-bound, bound), requires_grad=False)
                self.register_parameter(name + rev, p)

    def flatten_parameters(self):   # API compatibility (TransducerModel.py:54-55); nothing to flatten
        return None


class _LinearParameters(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        bound = 1.0 / math.sqrt(in_features)
        self.weight = nn.Parameter(torch.empty(out_features, in_features).uniform_(-bound, bound), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(out_features).uniform_(-bound, bound), requires_grad=False)


class TransducerGRU(nn.Module):
    def __init__(self, image_channels, image_features, gru_layers, hidden_size, num_base_classes, num_rle_classes,
                 bidirectional=True):
        super(TransducerGRU, self).__init__()
        if gru_layers != 1 or hidden_size != 128 or not bidirectional:
            raise ValueError("helen_b200 supports the shipped HELEN configuration only: gru_layers=1, "
                             f"hidden_size=128, bidirectional (got {gru_layers}, {hidden_size}, {bidirectional})")
        self.hidden_size = hidden_size
        self.bidirectional = bidirectional
        self.num_layers = gru_layers
        self.num_base_classes = num_base_classes
        self.num_rle_classes = num_rle_classes
        self.gru_encoder = _GRUParameters(image_features, hidden_size)
        self.gru_decoder = _GRUParameters(2 * hidden_size, hidden_size)
        self.dense1_base = _LinearParameters(2 * hidden_size, num_base_classes)
        self.dense2_rle = _LinearParameters(2 * hidden_size, num_rle_classes)
        self._predictor = None
        self._predictor_key = None

    def _native(self, device):
        key = (device.index or 0,) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._predictor is None or self._predictor_key != key:
            if self._predictor is not None:
                self._predictor.close()
            self._predictor = WindowPredictor(self.state_dict(), device=device.index or 0)
            self._predictor_key = key
        return self._predictor

    def predictor(self, device=0):
        """The window-level fused predictor built from this model's current parameters."""
        return self._native(torch.device("cuda", device) if not isinstance(device, torch.device) else device)

    def forward(self, x, hidden):
        """x [B, W, F], hidden [B, 2, H] (CUDA tensors) -> base [B, W, 5], rle [B, W, 11], hidden [B, 2, H]."""
        if not x.is_cuda:
            raise RuntimeError("helen_b200.TransducerGRU.forward needs CUDA tensors; there is no CPU execution path")
        return self._native(x.device).forward_chunk(x, hidden)

    def init_hidden(self, batch_size, num_layers, bidirectional=True):
        num_directions = 2 if bidirectional else 1
        return torch.zeros(batch_size, num_directions * num_layers, self.hidden_size)
