"""-m gpu: gradients of the training path against the reference's own autograd (goldens made by
tests/golden/make_golden.py::gen_grads on CPU).  Metric: max |dg| / max |g_ref| per tensor (grads are
sums of many atomically-accumulated fp32 terms; the tolerance is 2e-4 of the tensor's scale)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


def scale_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def test_lookup_backward_matches_torch_grid_sample():
    """The two CUDA backward kernels against autograd of F.grid_sample on the same parameters."""
    import torch.nn.functional as F
    import localrf_b200 as L
    from gpu_helpers import AABB, field_kwargs
    torch.manual_seed(3)
    sc = dict(app_dim=27, density_shift=-5.0, distance_scale=25.0, rayMarch_weight_thres=1e-3,
              view_pe=0, fea_pe=0, featureC=128, step_ratio=0.5, fea2denseAct="softplus")
    m = L.TensorVMSplit("cuda", AABB.clone().cuda(), [18, 22, 26], **field_kwargs(sc))
    xyz = (torch.rand(700, 3, device="cuda") * 2.2 - 1.1).requires_grad_(True)   # some clipped
    gd = torch.randn(700, device="cuda"); ga = torch.randn(700, 27, device="cuda")

    def ref_features(kind):
        planes, lines = (m.density_plane, m.density_line) if kind == "d" else (m.app_plane, m.app_line)
        outs = []
        for i in range(3):
            cp = xyz[:, list(m.matMode[i])].view(1, -1, 1, 2)
            cl = torch.stack([torch.zeros_like(xyz[:, 0]), xyz[:, m.vecMode[i]]], -1).view(1, -1, 1, 2)
            p = F.grid_sample(planes[i], cp, align_corners=True, padding_mode="border").view(-1, xyz.shape[0])
            l = F.grid_sample(lines[i], cl, align_corners=True, padding_mode="border").view(-1, xyz.shape[0])
            outs.append(p * l)
        if kind == "d":
            return sum(o.sum(0) for o in outs)
        return m.basis_mat(torch.cat(outs).T)

    for kind, fn, g in (("d", m.compute_densityfeature, gd), ("a", m.compute_appfeature, ga)):
        params = [p for p in m.parameters() if p.requires_grad]
        m.zero_grad(); xyz.grad = None
        (ref_features(kind) * g).sum().backward()
        ref = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
        ref_xyz = xyz.grad.clone()
        m.zero_grad(); xyz.grad = None
        out = fn(xyz)
        (out * g).sum().backward()
        assert scale_err(xyz.grad.cpu(), ref_xyz.cpu()) < 2e-5, kind
        for n, p in m.named_parameters():
            if n in ref:
                assert p.grad is not None, n
                assert scale_err(p.grad.cpu(), ref[n].cpu()) < 2e-5, (kind, n)


def test_field_gradients_vs_reference():
    from gpu_helpers import module_from_golden
    g = load_golden("grads_field")
    m = module_from_golden(g)
    rays = torch.from_numpy(g["rays"]).cuda().requires_grad_(True)
    z = torch.from_numpy(g["z"]).cuda()
    rgb, depth = m(rays, is_train=True, z_vals=z)
    assert rel_err(rgb.detach().cpu().numpy(), g["out.rgb"]) < 1e-4
    assert rel_err(depth.detach().cpu().numpy(), g["out.depth"]) < 1e-4
    loss = (rgb * torch.from_numpy(g["c_rgb"]).cuda()).sum() + (depth * torch.from_numpy(g["c_depth"]).cuda()).sum()
    loss.backward()
    assert scale_err(rays.grad.cpu().numpy(), g["grad.rays"]) < 2e-4
    checked = 0
    for name, p in m.named_parameters():
        key = "grad." + name
        if key in g:
            assert p.grad is not None, name
            assert p.grad.shape == tuple(g[key].shape)
            assert scale_err(p.grad.cpu().numpy(), g[key]) < 2e-4, name
            checked += 1
    assert checked == 19      # 12 grids + basis + 3 weights + 3 biases


def test_scene_gradients_vs_reference():
    """LocalTensorfs train-mode call: gradients of poses, intrinsics, exposure and the newest field."""
    from gpu_helpers import local_from_golden
    g = load_golden("grads_local")
    lt = local_from_golden(g)
    z = torch.from_numpy(g["z"]).cuda()
    rf = lt.tensorfs[-1]
    rf.sample_table = lambda *a, **k: z
    ids = torch.from_numpy(g["ray_ids"]).cuda()
    v = torch.from_numpy(g["view_ids"]).cuda()
    rgb, depth, dirs, ij = lt(ids, v, int(g["W"]), int(g["H"]), is_train=True)
    assert rel_err(rgb.detach().cpu().numpy(), g["out.rgb"]) < 1e-4
    assert rel_err(depth.detach().cpu().numpy(), g["out.depth"]) < 1e-4
    loss = (rgb * torch.from_numpy(g["c_rgb"]).cuda()).sum() + (depth * torch.from_numpy(g["c_depth"]).cuda()).sum()
    loss.backward()
    checked = 0
    for name, p in lt.named_parameters():
        key = "grad." + name
        if key in g:
            assert p.grad is not None, name
            assert scale_err(p.grad.cpu().numpy(), g[key]) < 3e-4, name
            checked += 1
    assert checked >= 30


def test_eval_path_still_fused_under_no_grad(monkeypatch):
    """Under torch.no_grad() the fused kernel is used even for parameters that require grad."""
    from gpu_helpers import module_from_golden
    g = load_golden("grads_field")
    m = module_from_golden(g)
    called = []
    monkeypatch.setattr(m, "_forward_autograd", lambda *a, **k: called.append(1))
    with torch.no_grad():
        m(torch.from_numpy(g["rays"]).cuda())
    assert not called


def test_optimizer_step_like_train_py():
    """The reference's training-loop shape (train.py:349-437): sample -> forward -> L1 loss ->
    LocalTensorfs.optimizer_step (backward + Adam on field, poses, exposure); the loss goes down."""
    import bench
    import localrf_b200 as L
    torch.manual_seed(0)
    aabb = 2 * torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    lt = L.LocalTensorfs(camera_prior=None, fov=85.6, n_init_frames=4, n_overlap=30, WH=(64, 48),
                         n_iters_per_frame=600, n_iters_reg=100, lr_R_init=5e-3, lr_t_init=5e-4,
                         lr_i_init=0, lr_exposure_init=1e-3, rf_lr_init=2e-2, rf_lr_basis=1e-3,
                         lr_decay_target_ratio=0.1, N_voxel_list={}, update_AlphaMask_list=[],
                         lr_upsample_reset=True, device="cuda", aabb=aabb.cuda(), gridSize=[48, 48, 48],
                         **bench.field_kwargs()).cuda()
    lt.is_refining = True
    g = torch.Generator().manual_seed(1)
    views = torch.tensor([0, 1, 2, 3], device="cuda")
    px = torch.randint(0, 64 * 48, (4, 256), generator=g)
    ids = (px + torch.arange(4)[:, None] * 64 * 48).reshape(-1).cuda()
    target = torch.full((1024, 3), 0.25, device="cuda")
    before = lt.tensorfs[-1].density_plane[0].detach().clone()
    pose_before = lt.t_c2w[1].detach().clone()
    losses = []
    for it in range(12):
        rgb, depth, dirs, ij = lt(ids, views, 64, 48, is_train=True)
        loss = (rgb - target).abs().mean()
        lt.optimizer_step(loss, optimize_poses=True)
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0] * 0.9, losses
    assert not torch.equal(before, lt.tensorfs[-1].density_plane[0].detach())
    assert not torch.equal(pose_before, lt.t_c2w[1].detach())
    assert lt.tensorfs[-1].density_plane[0].is_contiguous(memory_format=torch.channels_last)
