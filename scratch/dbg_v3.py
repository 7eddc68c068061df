import sys, os, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from helpers import load_golden
from gpu_helpers import module_from_golden
g = load_golden("cfg1_64"); m = module_from_golden(g)
rays = torch.from_numpy(g["rays"]).cuda(); z = torch.from_numpy(g["eval.z"]).cuda()
n = int(os.environ.get("NRAYS", "512"))
for it in range(int(os.environ.get("ITERS", "2"))):
    with torch.no_grad():
        rgb, depth = m(rays[:n], z_vals=z)
    torch.cuda.synchronize()
    e = np.abs(rgb.cpu().numpy() - g["eval.rgb"][:n]).max(1)
    bad = np.nonzero(e > 1e-4)[0]
    print("iter", it, "n", n, "nprod", os.environ.get("LRF_NPROD"), "bad", len(bad), bad[:24])
