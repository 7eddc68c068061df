"""Small fused forward + backward (both shade kernels), the schedule-time kernels and the frame pipeline, for
compute-sanitizer runs:  compute-sanitizer --tool memcheck|synccheck python tools/sanitize_backward.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import load_golden
from gpu_helpers import module_from_golden
import localrf_b200 as L

g = load_golden("grads_field"); m = module_from_golden(g)
rays0 = torch.from_numpy(g["rays"]).cuda()
z = torch.from_numpy(g["z"]).cuda()
c_rgb, c_depth = torch.from_numpy(g["c_rgb"]).cuda(), torch.from_numpy(g["c_depth"]).cuda()
for tc in ("1", "0"):
    os.environ["LRF_BWD_TC"] = tc
    m.zero_grad()
    rays = rays0.clone().requires_grad_(True)
    rgb, depth = m(rays, is_train=True, z_vals=z)
    ((rgb * c_rgb).sum() + (depth * c_depth).sum()).backward()
    torch.cuda.synchronize()
    err = max(float((p.grad.cpu() - torch.from_numpy(g["grad." + k])).abs().max() / max(np.abs(g["grad." + k]).max(), 1e-12))
              for k, p in m.named_parameters() if "grad." + k in g)
    print(f"backward (LRF_BWD_TC={tc}): worst gradient error {err:.2e} of scale")
# schedule-time kernels
g2 = load_golden("sched_nc"); m2 = module_from_golden(g2)
v = m2.density_L1(); v.backward()
m2.getDenseAlpha((10, 12, 14)); m2.updateAlphaMask((10, 12, 14)); m2.upsample_volume_grid([30, 33, 41])
r = torch.from_numpy(g2["sr.rays"]).cuda(); m2.sample_ray(r[:, :3], r[:, 3:], is_train=False, N_samples=40)
torch.cuda.synchronize(); print("schedule kernels ok", float(v))
# frame pipeline (zero-copy pinned outputs)
import bench
lt = bench.build_scene("cuda", 32)
pipe = L.FramePipeline(lt, 40, 32, n_buffers=2)
out = list(pipe.render([torch.tensor([0], device="cuda")] * 3))
print("pipeline frames", len(out))
