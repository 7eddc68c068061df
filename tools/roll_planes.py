"""Experiment (DESIGN.md par. 10, code size): time cfg2 batches with an alternative build of the library
(e.g. -DLRF_ROLL_PLANES=1) and dump one batch's pixels for comparison with the default build.
    python tools/roll_planes.py <path to .so> <tag>
The variant of tools/gpu_call19.sh was built with
    cd localrf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
        -shared -DLRF_ROLL_PLANES=1 -o liblrf_b200_roll.so lrf_render.cu lrf_aux.cu lrf_grad.cu lrf_backward.cu \
        lrf_sched.cu lrf_abi.cu"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from localrf_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1]); _lib._stale = lambda: False     # load exactly this binary
tag = sys.argv[2]
import localrf_b200 as L
import bench
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
wl = bench.Workload("cfg2")
lt = wl.build(L.LocalTensorfs, quiet=True).to(dev)
kw = wl.call_kwargs(lt, dev)
ids, views = wl.batches()
ids = ids.to(dev); view = views[0].to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(6):
    lt(ids[(i * 37) % ids.shape[0]], view, 800, 800, **kw); flush.zero_()
torch.cuda.synchronize()
K = 60
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
for i in range(K):
    flush.zero_(); flush.zero_()
    ev[i][0].record(); lt(ids[((6 + i) * 37) % ids.shape[0]], view, 800, 800, **kw); ev[i][1].record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
rgb, depth, _, _ = lt(ids[77], view, 800, 800, **kw)
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"roll_{tag}_b77.npy"), torch.cat([rgb, depth[:, None]], 1).cpu().numpy())
print(f"{tag}: mean {sum(ms) / K:.4f} ms  median {ms[K // 2]:.4f} ms  min {ms[0]:.4f} ms  -> {4096 / (sum(ms) / K) / 1e3:.2f} M rays/s")
