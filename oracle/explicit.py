"""Explicit (gate-by-gate) numpy restatement of the HELEN predict hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

What is restated, with the reference file:line each function follows
(paths relative to the reference checkout, kishwarshafin/helen @ a075e9f):

* the model arithmetic  -- ``helen/modules/python/models/TransducerModel.py:60-79``
  (``forward``): transpose hidden, bidirectional GRU encoder, bidirectional GRU decoder
  seeded with the encoder's final state, two affine heads, transpose back.  The GRU cell
  itself lives in third-party ``torch`` (``nn.GRU``; unpinned in the reference's
  ``requirements.txt:5``).  Its published definition (gate rows ordered r, z, n;
  ``b_hn`` inside the ``r *`` product) is restated in :func:`_gru_direction`.
* the sliding-window loop -- ``helen/modules/python/models/predict.py:90-154``
  (GPU twin ``predict_gpu.py:97-159``): float cast of the uint8 image, zero hidden per
  window, chunks of TRAIN_WINDOW columns every WINDOW_JUMP columns
  (``helen/modules/python/Options.py:25-26``), softmax per chunk, zero-pad + add
  (``predict.py:131-151``), first-index argmax (``predict.py:153-154``).

This implementation deliberately does not use ``torch.nn.GRU`` so that it pins gate
order / bias placement independently of the port in ``oracle/torch_port.py``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

HIDDEN = 128          # Options.py:28 HIDDEN_SIZE
N_BASE = 5            # Options.py:20 TOTAL_BASE_LABELS
N_RLE = 11            # Options.py:21 TOTAL_RLE_LABELS
SEQ_LENGTH = 1000     # Options.py:16
TRAIN_WINDOW = 100    # Options.py:25
WINDOW_JUMP = 50      # Options.py:26

_GRU_SUFFIXES = ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")
STATE_DICT_KEYS: Tuple[str, ...] = tuple(
    f"{layer}.{s}{rev}"
    for layer in ("gru_encoder", "gru_decoder")
    for rev in ("", "_reverse")
    for s in _GRU_SUFFIXES
) + ("dense1_base.weight", "dense1_base.bias", "dense2_rle.weight", "dense2_rle.bias")


def state_dict_shapes(image_features: int, hidden: int = HIDDEN) -> Dict[str, Tuple[int, ...]]:
    """Shapes of the 20 parameter tensors (TransducerModel.py:43-58)."""
    g = 3 * hidden
    shapes: Dict[str, Tuple[int, ...]] = {}
    for layer, k in (("gru_encoder", image_features), ("gru_decoder", 2 * hidden)):
        for rev in ("", "_reverse"):
            shapes[f"{layer}.weight_ih_l0{rev}"] = (g, k)
            shapes[f"{layer}.weight_hh_l0{rev}"] = (g, hidden)
            shapes[f"{layer}.bias_ih_l0{rev}"] = (g,)
            shapes[f"{layer}.bias_hh_l0{rev}"] = (g,)
    shapes["dense1_base.weight"] = (N_BASE, 2 * hidden)
    shapes["dense1_base.bias"] = (N_BASE,)
    shapes["dense2_rle.weight"] = (N_RLE, 2 * hidden)
    shapes["dense2_rle.bias"] = (N_RLE,)
    return shapes


@dataclass
class OracleWeights:
    """Parameter set in state_dict layout, as numpy arrays of one dtype."""

    tensors: Dict[str, np.ndarray]
    dtype: np.dtype

    @classmethod
    def from_state_dict(cls, state_dict, dtype=np.float32) -> "OracleWeights":
        tensors = {}
        for key, value in state_dict.items():
            # ModelHander.py:69-74: checkpoints written under DataParallel/DDP carry a
            # leading "module." that the loader strips.
            name = key[7:] if key.startswith("module.") else key
            arr = value.detach().cpu().numpy() if hasattr(value, "detach") else np.asarray(value)
            tensors[name] = np.ascontiguousarray(arr, dtype=dtype)
        missing = [k for k in STATE_DICT_KEYS if k not in tensors]
        if missing:
            raise KeyError(f"state_dict is missing {missing}")
        return cls(tensors=tensors, dtype=np.dtype(dtype))

    @property
    def image_features(self) -> int:
        return int(self.tensors["gru_encoder.weight_ih_l0"].shape[1])

    @property
    def hidden(self) -> int:
        return int(self.tensors["gru_encoder.weight_hh_l0"].shape[1])


def _sigmoid(x: np.ndarray) -> np.ndarray:
    # overflow-free form of 1 / (1 + exp(-x))
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def _gru_direction(x: np.ndarray, h0: np.ndarray, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """One direction of a one-layer GRU over a [B, W, K] sequence.

    torch.nn.GRU definition (called at TransducerModel.py:70,72): rows of weight_ih /
    weight_hh are ordered (r, z, n);
        r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)
        z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
        n = tanh(W_in x + b_in + r * (W_hn h + b_hn))
        h' = (1 - z) * n + z * h
    The reverse direction walks t = W-1 .. 0 and its state after consuming column t is
    the output at column t.
    """
    batch, width, _ = x.shape
    hid = h0.shape[1]
    y = np.empty((batch, width, hid), dtype=x.dtype)
    h = h0
    gi_all = x @ w_ih.T + b_ih            # [B, W, 3H]
    order = range(width - 1, -1, -1) if reverse else range(width)
    for t in order:
        gi = gi_all[:, t]
        gh = h @ w_hh.T + b_hh
        r = _sigmoid(gi[:, :hid] + gh[:, :hid])
        z = _sigmoid(gi[:, hid:2 * hid] + gh[:, hid:2 * hid])
        n = np.tanh(gi[:, 2 * hid:] + r * gh[:, 2 * hid:])
        h = (1.0 - z) * n + z * h
        y[:, t] = h
    return y, h


def _bigru(x, h0_pair, tensors, layer: str):
    """Bidirectional layer: output [B, W, 2H] = concat(fwd, bwd); h_n = (fwd, bwd)."""
    outs, finals = [], []
    for d, rev in enumerate(("", "_reverse")):
        y, hn = _gru_direction(
            x, h0_pair[d],
            tensors[f"{layer}.weight_ih_l0{rev}"], tensors[f"{layer}.weight_hh_l0{rev}"],
            tensors[f"{layer}.bias_ih_l0{rev}"], tensors[f"{layer}.bias_hh_l0{rev}"],
            reverse=bool(d),
        )
        outs.append(y)
        finals.append(hn)
    return np.concatenate(outs, axis=2), finals


def forward_chunk(weights: OracleWeights, x: np.ndarray, hidden: np.ndarray):
    """TransducerGRU.forward (TransducerModel.py:60-79).

    x [B, W, F], hidden [B, 2, H]  ->  base [B, W, 5], rle [B, W, 11], hidden [B, 2, H].
    """
    t = weights.tensors
    x = np.asarray(x, dtype=weights.dtype)
    hidden = np.asarray(hidden, dtype=weights.dtype)
    h0 = [hidden[:, 0], hidden[:, 1]]                       # :68 transpose(0,1): [0]=fwd, [1]=bwd
    y1, h1 = _bigru(x, h0, t, "gru_encoder")                # :70
    y2, h2 = _bigru(y1, h1, t, "gru_decoder")               # :72 decoder h0 = encoder h_n
    base = y2 @ t["dense1_base.weight"].T + t["dense1_base.bias"]   # :75
    rle = y2 @ t["dense2_rle.weight"].T + t["dense2_rle.bias"]      # :76
    hidden_out = np.stack(h2, axis=1)                       # :78 transpose back -> [B, 2, H]
    return base, rle, hidden_out


def chunk_starts(seq_len: int, window: int = TRAIN_WINDOW, jump: int = WINDOW_JUMP) -> List[int]:
    """Chunk start columns of predict.py:112-115: range(0, T, J), stop at i + W > T."""
    starts = []
    for i in range(0, seq_len, jump):
        if i + window > seq_len:
            break
        starts.append(i)
    return starts


def _softmax_last(x: np.ndarray) -> np.ndarray:
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


def predict_windows(weights: OracleWeights, images_u8: np.ndarray,
                    window: int = TRAIN_WINDOW, jump: int = WINDOW_JUMP):
    """Whole-window semantics of predict.py:90-154 for images_u8 [B, T, F] uint8.

    Returns dict(base_prob [B,T,5], rle_prob [B,T,11], base_label u8 [B,T],
    rle_label u8 [B,T], hidden [B,2,H] final carry).
    """
    images_u8 = np.asarray(images_u8)
    if images_u8.dtype != np.uint8:
        raise TypeError("images must be uint8 (dataloader_predict.py:69)")
    batch, seq_len, _ = images_u8.shape
    images = images_u8.astype(weights.dtype)                 # predict.py:92 FloatTensor cast
    hidden = np.zeros((batch, 2, weights.hidden), dtype=weights.dtype)   # :94
    p_base = np.zeros((batch, seq_len, N_BASE), dtype=weights.dtype)     # :104
    p_rle = np.zeros((batch, seq_len, N_RLE), dtype=weights.dtype)       # :105
    for i in chunk_starts(seq_len, window, jump):            # :112-116
        base, rle, hidden = forward_chunk(weights, images[:, i:i + window], hidden)   # :123
        # :131-151 softmax(dim=2) -> zero-pad to T -> add  ==  add into the slice
        p_base[:, i:i + window] += _softmax_last(base)
        p_rle[:, i:i + window] += _softmax_last(rle)
    # :153-154 torch.max returns the first maximal index; np.argmax does too.
    return {
        "base_prob": p_base,
        "rle_prob": p_rle,
        "base_label": p_base.argmax(axis=2).astype(np.uint8),
        "rle_label": p_rle.argmax(axis=2).astype(np.uint8),
        "hidden": hidden,
    }


def top2_margin(prob: np.ndarray) -> np.ndarray:
    """Per-position gap between the largest and second-largest accumulated score."""
    part = np.sort(prob, axis=-1)
    return part[..., -1] - part[..., -2]
