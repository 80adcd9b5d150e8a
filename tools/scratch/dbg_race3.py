import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import random_state_dict
from helen_b200.predictor import WindowPredictor
batch, seq, features = 45, 250, 10
sd = random_state_dict(features, seed=5)
gen = torch.Generator().manual_seed(77)
images = torch.randint(0, 256, (batch, seq, features), dtype=torch.uint8, generator=gen).cuda()
ref_pred = WindowPredictor(sd, device=0); ref_pred.set_engine("fp32")
ref = [t.cpu().numpy() for t in ref_pred.predict(images, return_probs=True)]; ref_pred.close()
for env in ({"HB_WINDOWS_PER_CTA": "16"}, {}, {"HB_GATE_WARPS": "16"}, {"HB_WINDOWS_PER_CTA": "16", "HB_GATE_WARPS": "16"}, {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "32", "HB_NO_PINGPONG": "1"}, {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "32", "HB_NO_PINGPONG": "1", "HB_NO_PDL": "1"},
            {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "16", "HB_NO_PINGPONG": "1"}, {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "8"}, {"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "32"}):
    for k in ("HB_WINDOWS_PER_CTA", "HB_GATE_WARPS", "HB_NO_CHUNKLOOP", "HB_NO_PINGPONG", "HB_NO_PDL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    pred = WindowPredictor(sd, device=0)
    fails, detail = 0, []
    for rep in range(300):
        got = pred.predict(images, return_probs=True)
        e = np.abs(got[3].cpu().numpy() - ref[3]).max(axis=2)
        bad = np.argwhere(e > 5e-6)
        if len(bad):
            fails += 1
            if len(detail) < 3:
                detail.append((rep, len(bad), sorted(set(int(w) for w in bad[:, 0]))[:12], int(bad[:, 1].min()), int(bad[:, 1].max()), float(e.max())))
    print(env, "fails %d / 300" % fails, detail, flush=True)
    pred.close()
