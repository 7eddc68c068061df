"""Launch time of small eval batches (what one rank renders under strong scaling), 300^3, coherent rays."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.set_grad_enabled(False)
lt = bench.build_scene(torch.device("cuda"), 300)
v = torch.tensor([0], device="cuda")
ids = torch.arange(640000, device="cuda")
for n in (256, 512, 1024, 2048, 4096, 8192):
    b = ids[320000:320000 + n]
    for _ in range(5):
        lt(b, v, 800, 800, is_train=False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
    for a, c in ev:
        a.record(); lt(b, v, 800, 800, is_train=False); c.record()
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)[len(ev) // 2]
    print(f"{n} rays: {t * 1e3:.1f} us per launch, {n / t / 1e3:.2f} M rays/s")
