import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import random_state_dict
from helen_b200.predictor import WindowPredictor

def run(variant_env, big_first):
    for k in ("HB_WINDOWS_PER_CTA", "HB_GATE_WARPS"):
        os.environ.pop(k, None)
    if big_first:
        sd = random_state_dict(10, seed=1)
        p = WindowPredictor(sd, device=0)
        imgs = torch.randint(0, 256, (big_first, 1000, 10), dtype=torch.uint8).cuda()
        p.predict(imgs); p.predict(imgs, return_probs=True); torch.cuda.synchronize(); p.close()
        del imgs
    batch, seq, features = 45, 250, 10
    sd = random_state_dict(features, seed=5)
    gen = torch.Generator().manual_seed(77)
    images = torch.randint(0, 256, (batch, seq, features), dtype=torch.uint8, generator=gen).cuda()
    ref_pred = WindowPredictor(sd, device=0); ref_pred.set_engine("fp32")
    ref = [t.cpu().numpy() for t in ref_pred.predict(images, return_probs=True)]; ref_pred.close()
    os.environ.update(variant_env)
    pred = WindowPredictor(sd, device=0)
    for rep in range(4):
        got = [t.cpu().numpy() for t in pred.predict(images, return_probs=True)]
        e = np.abs(got[3] - ref[3]).max(axis=2)          # [B, T]
        bad = np.argwhere(e > 5e-6)
        print(variant_env, "big", big_first, "rep", rep, "max err %.2e" % e.max(), "bad positions", len(bad),
              "windows", sorted(set(bad[:, 0]))[:20], "cols", (bad[:, 1].min(), bad[:, 1].max()) if len(bad) else None, pred.last_launch_plan())
    pred.close()

for big in (0, 2048, 512, 300):
    run({"HB_WINDOWS_PER_CTA": "16"}, big)
    run({"HB_WINDOWS_PER_CTA": "16", "HB_GATE_WARPS": "16"}, big)
    run({}, big)
