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
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
for env in ({"HB_WINDOWS_PER_CTA": "16"}, {"HB_WINDOWS_PER_CTA": "16", "HB_SYNC_MODE": "1"}, {"HB_WINDOWS_PER_CTA": "16", "HB_SYNC_MODE": "2"},
            {"HB_WINDOWS_PER_CTA": "16", "HB_SYNC_MODE": "4"}, {"HB_WINDOWS_PER_CTA": "16", "HB_SYNC_MODE": "7"},
            {"HB_WINDOWS_PER_CTA": "16", "HB_GATE_WARPS": "16"}, {"HB_WINDOWS_PER_CTA": "16", "HB_NO_PIXEL_JOBS": "1"}, {}):
    for k in ("HB_WINDOWS_PER_CTA", "HB_GATE_WARPS", "HB_SYNC_MODE", "HB_NO_PIXEL_JOBS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    pred = WindowPredictor(sd, device=0)
    fails, detail = 0, []
    for rep in range(reps):
        got = pred.predict(images, return_probs=True)
        e = (got[3].cpu().numpy() - ref[3])
        e = np.abs(e).max(axis=2)
        bad = np.argwhere(e > 5e-6)
        if len(bad):
            fails += 1
            if len(detail) < 4:
                detail.append((rep, len(bad), sorted(set(int(w) for w in bad[:, 0])), int(bad[:, 1].min()), int(bad[:, 1].max()), float(e.max())))
    print(env, "fails %d / %d" % (fails, reps), detail, flush=True)
    pred.close()
