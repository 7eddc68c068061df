import sys, time, torch, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from helpers import rel_err
lt = bench.build_scene(torch.device("cuda"), 640)
rf = lt.tensorfs[0]
print("S", rf.sample_table(False,-1,torch.device("cuda")).numel(), "nSamples", rf.nSamples)
v = torch.tensor([0], device="cuda")
ids = torch.arange(80*4096, 81*4096, device="cuda")
for thr in (0.0, 0.5):
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        rgb, depth, _, _ = lt(ids, v, 800, 800, is_train=False, floater_thresh=thr, stats=stats)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(10): lt(ids, v, 800, 800, is_train=False, floater_thresh=thr)
    torch.cuda.synchronize(); dt = (time.perf_counter()-t0)/10
    sel = np.arange(0, 4096, 97)[:40]
    field = bench.oracle_field(lt)
    from oracle import oracle as orc
    z = orc.sample_table(field.n_samples())
    focal = float(lt.focal(800).detach().cpu()); cx, cy = [float(x) for x in lt.center(800,800).detach().cpu()]
    c2w = lt.get_cam2world([0]).detach().cpu().numpy()
    expo = torch.stack(list(lt.exposure))[[0]].detach().cpu().numpy()
    ref = orc.local_forward([field],[z], ids.cpu().numpy()[sel], 800, 800, False, focal, cx, cy, c2w, np.zeros((1,3),np.float32), np.ones((1,1),np.float32), exposure=expo, floater_thresh=thr)
    print("floater", thr, "stats/ray", [s/4096 for s in stats.tolist()], "ms/batch", dt*1e3, "rays/s", 4096/dt,
          "rgb err", rel_err(rgb.cpu().numpy()[sel], ref["rgb"]), "depth err", rel_err(depth.cpu().numpy()[sel], ref["depth"]))
