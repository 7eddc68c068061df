import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
lt = bench.build_scene(torch.device("cuda"), 300)
lt.is_refining = True
ids = torch.randint(0, 640000, (4096,), device="cuda")
v = torch.tensor([0], device="cuda")
target = torch.rand(4096, 3, device="cuda")
def step():
    rgb, depth, _, _ = lt(ids, v, 800, 800, is_train=True)
    loss = (rgb - target).abs().mean() + 1e-3 * depth.mean()
    lt.optimizer_step(loss, optimize_poses=True)
for path in ("fused", "composed"):
    os.environ["LRF_TRAIN_PATH"] = path          # fused: lrf_render + lrf_render_backward
    torch.cuda.reset_peak_memory_stats()
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"train step 300^3, 4096 rays [{path}]: {dt*1e3:.2f} ms/iter  ({4096/dt/1e6:.3f} M rays/s)  peak mem {torch.cuda.max_memory_allocated()/1e9:.2f} GB")
# backward alone (CUDA events around autograd.backward of the fused node)
os.environ["LRF_TRAIN_PATH"] = "fused"
rf = lt.tensorfs[0]
rays = torch.cat([(torch.rand(4096, 3, device="cuda") - 0.5) * 0.6,
                  torch.nn.functional.normalize(torch.randn(4096, 3, device="cuda"), dim=-1)], -1).requires_grad_(True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ts = []
for it in range(8):
    rf.zero_grad()
    ev[0].record(); rgb, depth = rf(rays, is_train=True); ev[1].record()
    (rgb.sum() + depth.sum()).backward(); ev[2].record(); torch.cuda.synchronize()
    ts.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
print("field forward / backward ms (last 5):", [(round(a, 3), round(b, 3)) for a, b in ts[-5:]])
