#!/usr/bin/env python
"""Run-to-run equality of full-size batches: python tools/gpu_repeat.py [--batch 512] [--reps 200] (GPU box).
Every repetition of the same batch must give the same labels as the first; prints the plan the library chose."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--reps", type=int, default=200)
ap.add_argument("--features", type=int, default=10)
args = ap.parse_args()
from helen_b200.predictor import WindowPredictor
from helen_b200.models.TransducerModel import TransducerGRU

torch.manual_seed(0)
sd = TransducerGRU(1, args.features, 1, 128, 5, 11).state_dict()
pred = WindowPredictor(sd, device=0)
images = torch.randint(0, 256, (args.batch, 1000, args.features), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).cuda()
base0, rle0 = pred.predict(images)
torch.cuda.synchronize()
bad = 0
for rep in range(args.reps):
    b, r = pred.predict(images)
    if not (torch.equal(b, base0) and torch.equal(r, rle0)):
        bad += 1
print("batch %d: plan %s, %d of %d repetitions differ from the first" % (args.batch, pred.last_launch_plan(), bad, args.reps))
sys.exit(1 if bad else 0)
