"""Small end-to-end render (cfg1 golden, 512 rays, incl. a 3-field LocalTensorfs call) for
compute-sanitizer runs:  compute-sanitizer --tool memcheck|synccheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import load_golden, rel_err
from gpu_helpers import module_from_golden, local_from_golden
g = load_golden("cfg1_64"); m = module_from_golden(g)
rays = torch.from_numpy(g["rays"]).cuda(); z = torch.from_numpy(g["eval.z"]).cuda()
with torch.no_grad():
    for thr in (0.0, 0.5):
        rgb, depth = m(rays, z_vals=z, floater_thresh=thr, return_weights=True)
torch.cuda.synchronize()
print("field: rgb err", rel_err(rgb.cpu().numpy(), g["eval_floater.rgb"]))
g2 = load_golden("local3"); lt = local_from_golden(g2)
with torch.no_grad():
    out = lt(torch.from_numpy(g2["ray_ids"]).cuda(), torch.tensor([2]).cuda(), int(g2["W"]), int(g2["H"]),
             is_train=False, blending_weights=torch.from_numpy(g2["blend3"]).cuda())
torch.cuda.synchronize()
print("scene: rgb err", rel_err(out[0].cpu().numpy(), g2["blend3.rgb"]))
