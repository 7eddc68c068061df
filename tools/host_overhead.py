"""Host-side cost of one eval call of LocalTensorfs.forward (plan fast path), measured with a tiny batch so
the GPU is never the bottleneck: calls per second without synchronising, and a cProfile of the call."""
import cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.set_grad_enabled(False)
lt = bench.build_scene(torch.device("cuda"), 64)
v = torch.tensor([0], device="cuda")
ids = torch.arange(64, dtype=torch.int64, device="cuda")
ids_pin = torch.arange(64, dtype=torch.int64).pin_memory()
rgb_h, d_h = torch.empty(64, 3).pin_memory(), torch.empty(64).pin_memory()
for name, call in (("device ids", lambda: lt(ids, v, 800, 800, is_train=False)),
                   ("pinned ids + pinned out", lambda: lt(ids_pin, v, 800, 800, is_train=False, out=(rgb_h, d_h)))):
    for _ in range(50): call()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2000): call()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name}: host {1e6 * (t1 - t0) / 2000:.1f} us per call (enqueue only), {1e6 * (t2 - t0) / 2000:.1f} us incl. drain")
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): lt(ids, v, 800, 800, is_train=False)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
