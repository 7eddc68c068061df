#!/usr/bin/env python
"""bench.py -- rays/s of the fused render path at a 300^3 VM grid, 4096-ray batches (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: one TensorVMSplit at 300^3 built with the reference's constructor
defaults and torch.manual_seed(0) ("distribution A", random init), an 800x800 pinhole frame at
fov 85.6, identity pose, cut into 157 batches of 4096 rays.  One STEP = one 4096-ray batch
(ray generation + march + shading + composite = one fused launch).  Steps walk through the frame's
batches, so consecutive steps use different rays.

  value     rays/s with the batch's ray ids already in HBM; each step timed with CUDA events on the
            launching stream, a 256 MiB write between steps flushes the 126 MB L2.
  e2e       the same step through the public API (LocalTensorfs.forward) from pinned HOST ray ids,
            host->device copy, launch, device->host copy of rgb+depth, host sync -- every step.
  roofline  algorithmic gather bytes (576 B per density sample, 1728 B per shaded sample, 40 B per
            ray; samples counted exactly by the kernel) / event-timed launch duration, against the
            measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the CPU oracle port (oracle/, plain C + OpenMP, all host cores)
            on a bounded sample of the same workload.

N > 1 (torchrun, one rank per GPU): the field is replicated, every rank renders its own 4096-ray
batch per step (global batch N*4096, weak scaling) and one NCCL all-gather of the rendered
[4096,4] pixels (rgb+depth) closes the step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 300
IMG_W = IMG_H = 800
BATCH = 4096
FOV = 85.6
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print (NCCL's version banner, torchrun notices) goes to stderr; only the
    ONE JSON line reaches the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def field_kwargs():
    return dict(density_n_comp=[8, 8, 8], appearance_n_comp=[24, 24, 24], app_dim=27,
                shadingMode="MLP_Fea_late_view", near_far=[0.1, 1e3], density_shift=-5,
                alphaMask_thres=1e-4, distance_scale=25, rayMarch_weight_thres=1e-3, pos_pe=0,
                view_pe=0, fea_pe=0, featureC=128, step_ratio=0.5, fea2denseAct="softplus")


def build_scene(device, grid=GRID):
    """LocalTensorfs with one 300^3 field, reference constructor defaults, seed 0."""
    import localrf_b200 as L
    torch.manual_seed(0)
    aabb = 2 * torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    lt = L.LocalTensorfs(camera_prior=None, fov=FOV, n_init_frames=1, n_overlap=30,
                         WH=(IMG_W, IMG_H), n_iters_per_frame=600, n_iters_reg=100, lr_R_init=5e-3,
                         lr_t_init=5e-4, lr_i_init=0, lr_exposure_init=1e-3, rf_lr_init=2e-2,
                         rf_lr_basis=1e-3, lr_decay_target_ratio=0.1, N_voxel_list={},
                         update_AlphaMask_list=[], lr_upsample_reset=True, device="cpu",
                         aabb=aabb, gridSize=[grid] * 3, **field_kwargs())
    return lt.to(device) if device != "cpu" else lt


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md).  NVML (pynvml) is
    polled every ~2 ms because the timed region is only tens of milliseconds long; nvidia-smi
    (one sample per ~100 ms) is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(
            n, "nvmlDeviceGetCurrentClocksEventReasons") else int(
            n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        flags = [bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4)]  # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
        self.samples.append([mhz, self.max_mhz] + ["Active" if f else "Not Active" for f in flags])

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                self.nvml = None
            time.sleep(0.02)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(str(s[2 + i]).lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def frame_ids():
    return torch.arange(IMG_W * IMG_H, dtype=torch.int64)


def oracle_field(lt):
    from oracle import oracle as orc
    rf = lt.tensorfs[0]
    fd = {k: v.detach().cpu().numpy() for k, v in rf.state_dict().items()}
    kw = rf.get_kwargs()
    for k in ("density_shift", "distance_scale", "rayMarch_weight_thres", "fea_pe", "view_pe",
              "featureC", "app_dim", "step_ratio", "fea2denseAct", "gridSize"):
        fd[k] = kw[k]
    return orc.Field(fd)


def oracle_batch(lt, field, ids, n_threads=None):
    """One batch through the CPU oracle's LocalTensorfs.forward restatement, on ALL host cores
    (explicit thread count: torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    from oracle import oracle as orc
    z = orc.sample_table(field.n_samples())
    focal = float(lt.focal(IMG_W).detach().cpu())
    cx, cy = [float(v) for v in lt.center(IMG_W, IMG_H).detach().cpu()]
    c2w = lt.get_cam2world(torch.tensor([0])).detach().cpu().numpy()
    expo = torch.stack(list(lt.exposure))[[0]].detach().cpu().numpy()
    return orc.local_forward([field], [z], ids, IMG_W, IMG_H, False, focal, cx, cy, c2w,
                             np.zeros((1, 3), np.float32), np.ones((1, 1), np.float32),
                             exposure=expo, n_threads=n_threads)


def cpu_baseline(lt, budget_s=12.0):
    """Oracle port on the host cores, bounded sample of the same workload."""
    field = oracle_field(lt)
    ids = frame_ids().numpy()
    mid = (IMG_W * IMG_H // 2 // BATCH) * BATCH          # a batch from the middle of the frame
    oracle_batch(lt, field, ids[mid:mid + 256])          # warm-up (page-in, thread pool)
    t0 = time.perf_counter()
    oracle_batch(lt, field, ids[mid:mid + 1024])
    t1 = time.perf_counter() - t0
    n = int(min(max(1024 * budget_s / max(t1, 1e-6), 1024), 16 * BATCH))
    n = (n // 1024) * 1024
    t0 = time.perf_counter()
    oracle_batch(lt, field, ids[mid:mid + n])
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} consecutive rays of the 800x800 frame (from ray {mid}), oracle/lrf_oracle.c "
                      f"with OpenMP over {os.cpu_count()} host threads, {dt:.2f} s"}


def run_reference(args):
    """--impl reference: the CPU oracle port on this box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lt = build_scene("cpu")
    field = oracle_field(lt)
    ids = frame_ids().numpy()
    t0 = time.perf_counter()
    oracle_batch(lt, field, ids[:256])
    oracle_batch(lt, field, ids[320000:320512])
    rate = 512 / max(time.perf_counter() - t0, 1e-6) * 1.2
    budget = 150.0 / (args.steps + args.warmup)
    n = int(min(max(rate * budget, 64), BATCH))
    n_batches = IMG_W * IMG_H // BATCH
    def step(i):
        lo = (i * 37 % n_batches) * BATCH
        oracle_batch(lt, field, ids[lo:lo + n])
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    emit({
        "impl": "reference", "metric": "rays/sec at 300^3 VM grid, 4096-ray batch", "value": value,
        "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg2: TensorVMSplit {GRID}^3 (seed 0, random init), 800x800 frame, "
                               f"4096-ray batches; each step = the first {n} rays of a batch"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{n} rays per step, oracle/lrf_oracle.c, OpenMP {os.cpu_count()} threads"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=157)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grid", type=int, default=GRID)
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import ctypes as C
    import localrf_b200 as L
    from localrf_b200 import _lib
    from localrf_b200.tensorf import _ptr, _stream

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    lt = build_scene(dev, args.grid)
    rf = lt.tensorfs[0]
    ids_dev = frame_ids().to(dev)
    n_batches = IMG_W * IMG_H // BATCH     # 156 full batches (+ a ragged one the bench skips)

    # ---- device-resident step: direct C-ABI launch --------------------------------------------------
    view = torch.tensor([0], device=dev)
    z = rf.sample_table(False, -1, dev)
    fs, prep = rf.field_and_prepared(z)
    S = z.numel()
    c2w = lt.get_cam2world(view).detach().contiguous()
    intr = torch.cat([lt.focal(IMG_W).detach().reshape(1), lt.center(IMG_W, IMG_H).detach().reshape(2)]).contiguous()
    expo = torch.stack(list(lt.exposure))[view].detach().contiguous()
    blend = torch.ones(1, 1, device=dev)
    rgb = torch.empty(BATCH, 3, device=dev); depth = torch.empty(BATCH, device=dev)
    gathered = torch.empty(world * BATCH, 4, device=dev) if world > 1 else None
    pix = torch.empty(BATCH, 4, device=dev) if world > 1 else None
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    b = _lib.LrfBatch(); o = _lib.LrfOutputs()
    b.n_rays = BATCH; b.W, b.H = IMG_W, IMG_H; b.fov360 = 0
    b.intrinsics = intr.data_ptr(); b.cam2world = c2w.data_ptr(); b.n_views = 1
    b.blend = blend.data_ptr(); b.blend_stride = 1; b.exposure = expo.data_ptr()
    b.accumulate = 0; b.finalize = 1; b.white_bg = 1; b.floater_thresh = 0.0
    o.rgb, o.depth, o.stats = rgb.data_ptr(), depth.data_ptr(), stats.data_ptr()
    if world > 1:
        o.pix = pix.data_ptr()          # interleaved (r,g,b,depth): one all-gather, no repack kernels
    lib = _lib.lib()
    stream = _stream(dev)

    def step(i):
        bi = ((i * world + rank) * 37) % n_batches        # a different batch of the frame each step
        b.ray_ids = ids_dev.data_ptr() + 8 * BATCH * bi
        _lib.check(lib.lrf_render(C.byref(fs), _ptr(prep), C.byref(b), C.byref(o), stream))
        if world > 1:
            dist.all_gather_into_tensor(gathered, pix)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(args.warmup):
        step(i); flush.zero_()
    torch.cuda.synchronize()
    stats.zero_()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the events)
        ev[i][0].record()
        step(args.warmup + i)
        ev[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    times_ms = [a.elapsed_time(c) for a, c in ev]
    total_ms = torch.tensor([sum(times_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    st = stats.cpu().tolist()
    value = world * BATCH * args.steps / (total_ms * 1e-3)

    # roofline of the fused kernel (rank 0's launches)
    bytes_total = st[0] * 576 + st[1] * 1728 + BATCH * args.steps * 40
    kern_ms = sum(times_ms) if world == 1 else None
    peak, peak_src = peaks()
    roofline = None
    if world == 1:
        achieved = bytes_total / (kern_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel from the
                    # committed `ncu --set full` capture (profiles/r1_v3_streaming.md): 4.08 MB + 0 B
                    "traffic": 4.08e6, "traffic_unit": "bytes per launch (ncu, profiles/r1_v3_streaming.md)",
                    "algorithmic_bytes_per_launch": bytes_total / args.steps,
                    "peak_source": peak_src,
                    "kernel": "lrf::render_kernel",
                    "bytes_per_ray": bytes_total / (BATCH * args.steps),
                    "density_samples_per_ray": st[0] / (BATCH * args.steps),
                    "app_samples_per_ray": st[1] / (BATCH * args.steps),
                    "note": "algorithmic gather bytes (every texel fetch counted); the 33 MB field is "
                            "L2-resident, so DRAM counters read far lower (profiles/)"}

    # ---- e2e: public API, host buffers, every step ---------------------------------------------------
    ids_host = frame_ids().pin_memory()
    rgb_host = torch.empty(BATCH, 3, pin_memory=True)
    depth_host = torch.empty(BATCH, pin_memory=True)
    e2e_steps = args.steps
    def e2e_step(i):
        bi = ((i * world + rank) * 37) % n_batches
        ids = ids_host[bi * BATCH:(bi + 1) * BATCH].to(dev, non_blocking=True)
        with torch.no_grad():
            r, d, _, _ = lt(ids, view, IMG_W, IMG_H, is_train=False, chunk=BATCH)
        rgb_host.copy_(r, non_blocking=True)
        depth_host.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
    for i in range(args.warmup):
        e2e_step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(args.warmup + i)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * e2e_steps / float(e2e_s)

    # ---- whole frame through the API (one fused launch for all 640k rays), informational ------------
    frame = None
    if world == 1:
        all_ids = ids_host
        def frame_call():
            ids = all_ids.to(dev, non_blocking=True)
            with torch.no_grad():
                r, d, _, _ = lt(ids, view, IMG_W, IMG_H, is_train=False, chunk=BATCH)
            r.cpu(); d.cpu()
        frame_call()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        frame_call()
        frame = {"rays_per_s": IMG_W * IMG_H / (time.perf_counter() - t0),
                 "what": "LocalTensorfs.forward on all 640000 rays of the frame incl. H2D/D2H"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(build_scene("cpu", args.grid))
    line = {
        "metric": "rays/sec at 300^3 VM grid, 4096-ray batch", "value": value, "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "dtype_note": "fp32 storage and arithmetic; the two dense MLP layers run on tcgen05 as three bf16 "
                      "products of hi/lo-split fp32 operands with fp32 TMEM accumulators (1e-7 of fp32 on rgb, "
                      "tests/test_gpu_parity.py::test_tensor_core_mlp_vs_torch_fp32)",
        "config": {"workload": f"cfg2: TensorVMSplit {args.grid}^3 (reference ctor, seed 0, random init), "
                               "800x800 pinhole frame fov 85.6 identity pose, 4096-ray batches; "
                               "step = one batch, steps walk the frame's batches",
                   "rays_per_batch": BATCH, "samples_per_ray": S, "l2": "256 MiB write between timed steps",
                   "parallelism": f"ray-batch data parallel x{world}" + (", all-gather [4096,4] per step" if world > 1 else "")},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": BATCH * 8,
                "d2h_bytes_per_step": BATCH * 16},
        "gpu_launches": args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "frame_api": frame,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
