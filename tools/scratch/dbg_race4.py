import os, sys, subprocess
code = r'''
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
for env in ({"HB_NO_CHUNKLOOP": "1", "HB_WINDOWS_PER_CTA": "32", "HB_NO_PINGPONG": "1"}, {"HB_WINDOWS_PER_CTA": "16"}):
    for k in ("HB_WINDOWS_PER_CTA", "HB_NO_CHUNKLOOP", "HB_NO_PINGPONG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    pred = WindowPredictor(sd, device=0)
    fails = 0
    for rep in range(200):
        got = pred.predict(images, return_probs=True)
        fails += np.abs(got[3].cpu().numpy() - ref[3]).max() > 5e-6
    print(os.environ.get("HB_LIB", "product").split("_")[-1], env, "fails %d / 200" % fails, flush=True)
    pred.close()
'''
for lib in (None, "noloadsdone", "scalargi", "upload2"):
    env = dict(os.environ)
    if lib:
        env["HB_LIB"] = os.path.join(os.getcwd(), "helen_b200/lib/libhelen_b200_%s.so" % lib)
    subprocess.run([sys.executable, "-c", code], env=env)
