#!/usr/bin/env python
"""Precision probe: how many consensus labels flip, and how far accumulated probabilities
move, when the GRU contractions run on split-precision tensor-core operands.

Emulates on CPU (numpy fp64 accumulate of rounded operands) the candidate operand formats
for the tcgen05 path; evidence for DESIGN.md's choice.  Not part of the product.

    python tests/precision_probe.py [F] [B]
"""
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import explicit as ex
from oracle import random_state_dict, OracleWeights


def bf16_round(a):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return t.to(torch.bfloat16).to(torch.float32).numpy()


def tf32_round(a):
    b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    b = (b + 0x1000) & 0xFFFFE000     # round-to-nearest on 13 dropped bits
    return b.view(np.float32)


def fp16_round(a):
    with np.errstate(over="raise"):
        return np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float32)


def split(a, rnd, terms):
    parts, rest = [], np.asarray(a, dtype=np.float32)
    for _ in range(terms):
        p = rnd(rest)
        parts.append(p)
        rest = (rest - p).astype(np.float32)
    return parts


class SplitMatmul:
    """x @ w.T with both operands split into `terms` parts, keeping products i+j < keep."""
    def __init__(self, rnd, terms, keep, x_scale=1.0, w_scale=1.0):
        self.rnd, self.terms, self.keep = rnd, terms, keep
        self.x_scale, self.w_scale = x_scale, w_scale     # powers of two: exact to apply / undo

    def __call__(self, x, w):
        xs_ = self.x_scale if np.abs(x).max() <= 1.0 else 1.0   # activations in (-1,1) get scaled; u8 pixels do not
        xs, ws = split(x * xs_, self.rnd, self.terms), split(w * self.w_scale, self.rnd, self.terms)
        inv = 1.0 / (xs_ * self.w_scale)
        acc = 0
        for i, xp in enumerate(xs):
            for j, wp in enumerate(ws):
                if i + j < self.keep:
                    acc = acc + (xp.astype(np.float64) @ wp.T.astype(np.float64))
        return (acc * (inv if hasattr(self, "w_scale") else 1.0)).astype(np.float32)


def run(weights, images, mm):
    t = weights.tensors
    B, T, F = images.shape
    H = 128
    hidden = [np.zeros((B, H), np.float32), np.zeros((B, H), np.float32)]
    pb = np.zeros((B, T, 5), np.float32); pr = np.zeros((B, T, 11), np.float32)
    def direction(x, h, layer, rev, reverse):
        w_ih, w_hh = t[f"{layer}.weight_ih_l0{rev}"], t[f"{layer}.weight_hh_l0{rev}"]
        b_ih, b_hh = t[f"{layer}.bias_ih_l0{rev}"], t[f"{layer}.bias_hh_l0{rev}"]
        W = x.shape[1]
        gi_all = mm(x.reshape(-1, x.shape[2]), w_ih).reshape(B, W, 3 * H) + b_ih
        y = np.empty((B, W, H), np.float32)
        for tt in (range(W - 1, -1, -1) if reverse else range(W)):
            gh = mm(h, w_hh) + b_hh
            gi = gi_all[:, tt]
            r = ex._sigmoid(gi[:, :H] + gh[:, :H]); z = ex._sigmoid(gi[:, H:2*H] + gh[:, H:2*H])
            n = np.tanh(gi[:, 2*H:] + r * gh[:, 2*H:])
            h = ((1 - z) * n + z * h).astype(np.float32)
            y[:, tt] = h
        return y, h
    for i in ex.chunk_starts(T):
        x = images[:, i:i+100].astype(np.float32)
        ys = []
        for d, rev in enumerate(("", "_reverse")):
            y, hidden[d] = direction(x, hidden[d], "gru_encoder", rev, bool(d)); ys.append(y)
        y1 = np.concatenate(ys, 2); ys = []
        for d, rev in enumerate(("", "_reverse")):
            y, hidden[d] = direction(y1, hidden[d], "gru_decoder", rev, bool(d)); ys.append(y)
        y2 = np.concatenate(ys, 2)
        hw = np.concatenate([t["dense1_base.weight"], t["dense2_rle.weight"]], 0)
        hb = np.concatenate([t["dense1_base.bias"], t["dense2_rle.bias"]], 0)
        lg = mm(y2.reshape(-1, 256), hw).reshape(B, 100, 16) + hb
        pb[:, i:i+100] += ex._softmax_last(lg[..., :5]); pr[:, i:i+100] += ex._softmax_last(lg[..., 5:])
    return pb, pr


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    sd = random_state_dict(F, seed=0)
    w32 = OracleWeights.from_state_dict(sd, np.float32)
    w64 = OracleWeights.from_state_dict(sd, np.float64)
    gen = torch.Generator().manual_seed(1)
    images = torch.randint(0, 256, (B, 1000, F), dtype=torch.uint8, generator=gen).numpy()
    ref64 = ex.predict_windows(w64, images)
    ref32 = ex.predict_windows(w32, images)
    print(f"F={F} B={B}: {B*1000} positions/head")
    def report(name, pb, pr):
        fb = (pb.argmax(2) != ref64["base_label"]).sum(); fr = (pr.argmax(2) != ref64["rle_label"]).sum()
        db = np.abs(pb - ref64["base_prob"]).max(); dr = np.abs(pr - ref64["rle_prob"]).max()
        print(f"  {name:24s} flips base {fb:4d} rle {fr:4d}   max|dP| base {db:.2e} rle {dr:.2e}")
    report("fp32 oracle", ref32["base_prob"], ref32["rle_prob"])
    for name, mm in [
        ("bf16 x1", SplitMatmul(bf16_round, 1, 1)),
        ("tf32 x1", SplitMatmul(tf32_round, 1, 1)),
        ("bf16 x3 (hh,hl,lh)", SplitMatmul(bf16_round, 2, 2)),
        ("bf16 x4 (all 2x2)", SplitMatmul(bf16_round, 2, 3)),
        ("bf16 x6 (3-way, i+j<3)", SplitMatmul(bf16_round, 3, 3)),
        ("tf32 x3", SplitMatmul(tf32_round, 2, 2)),
        ("fp16 x1", SplitMatmul(fp16_round, 1, 1)),
        ("fp16 x3 scaled 2^8/2^8", SplitMatmul(fp16_round, 2, 2, 256.0, 256.0)),
        ("fp16 x3 scaled 2^10/2^10", SplitMatmul(fp16_round, 2, 2, 1024.0, 1024.0)),
        ("fp16 x3 (hh,hl,lh)", SplitMatmul(fp16_round, 2, 2)),
    ]:
        pb, pr = run(w32, images, mm)
        report(name, pb, pr)


if __name__ == "__main__":
    main()
