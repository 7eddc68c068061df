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
for _ in range(3): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print(f"train step 300^3, 4096 rays: {dt*1e3:.2f} ms/iter  ({4096/dt/1e6:.3f} M rays/s)  peak mem {torch.cuda.max_memory_allocated()/1e9:.2f} GB")
