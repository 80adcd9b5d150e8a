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
os.environ["HB_WINDOWS_PER_CTA"] = "16"
pred = WindowPredictor(sd, device=0)
fails = 0
for rep in range(400):
    got = pred.predict(images, return_probs=True)
    e = np.abs(got[3].cpu().numpy() - ref[3]).max()
    fails += e > 5e-6
print(os.environ.get("HB_LIB", "product").split("_")[-1], "fails %d / 400" % fails, flush=True)
'''
for lib in (None, "scalargi", "inlineup", "latehfree", "oldgates"):
    env = dict(os.environ)
    if lib:
        env["HB_LIB"] = os.path.join(os.getcwd(), "helen_b200/lib/libhelen_b200_%s.so" % lib)
    subprocess.run([sys.executable, "-c", code], env=env)
