"""Tiny bf16-grid-storage run for compute-sanitizer (the third instantiation of the render kernel, the bf16
lookups, lrf_pack_bf16):  compute-sanitizer --tool memcheck python tools/sanitize_bf16.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_golden
from gpu_helpers import module_from_golden
m = module_from_golden(load_golden("opaque_32"))
with torch.no_grad():
    for plist in (m.density_plane, m.density_line, m.app_plane, m.app_line):
        for p in plist:
            p.copy_(p.to(torch.bfloat16).to(torch.float32))
g = torch.Generator().manual_seed(1)
rays = torch.cat([0.1 * torch.randn(192, 3, generator=g), torch.randn(192, 3, generator=g)], -1).cuda()
with torch.no_grad():
    a = m(rays, floater_thresh=0.5, return_weights=True)
    m.set_grid_storage("bf16")
    b = m(rays, floater_thresh=0.5, return_weights=True)
    c = m(rays)
torch.cuda.synchronize()
print("bf16 == fp32 storage:", torch.equal(a[0], b[0]), torch.equal(a[1], b[1]))
