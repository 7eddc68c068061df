"""A few fused forward+backward passes of one field at the BASELINE size (300^3, 4096 rays), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
lt = bench.build_scene(torch.device("cuda"), 300)
rf = lt.tensorfs[0]
g = torch.Generator().manual_seed(0)
o = (torch.rand(4096, 3, generator=g) - 0.5) * 0.6
d = torch.nn.functional.normalize(torch.randn(4096, 3, generator=g), dim=-1)
rays = torch.cat([o, d], -1).cuda().requires_grad_(True)
for it in range(int(os.environ.get("ITERS", "3"))):
    rf.zero_grad()
    rgb, depth = rf(rays, is_train=True)
    (rgb.sum() + depth.sum()).backward()
torch.cuda.synchronize()
print("done")
