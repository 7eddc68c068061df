import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torch.profiler import profile, ProfilerActivity
lt = bench.build_scene(torch.device("cuda"), 300); lt.is_refining = True
ids = torch.randint(0, 640000, (4096,), device="cuda"); v = torch.tensor([0], device="cuda")
target = torch.rand(4096, 3, device="cuda")
def step():
    rgb, depth, _, _ = lt(ids, v, 800, 800, is_train=True)
    loss = (rgb - target).abs().mean() + 1e-3 * depth.mean()
    lt.optimizer_step(loss, optimize_poses=True)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
