"""-m gpu: the schedule-time kernels and the AABB sampler against goldens the
unmodified reference produced on a non-cubic [20,24,28] grid (tests/golden/sched_nc.npz):
occupancy-mask rebuild, grid upsampling, density_L1 / TV losses with their gradients, sample_ray."""
import numpy as np
import pytest
import torch

from gpu_helpers import module_from_golden
from helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


class TVLoss(torch.nn.Module):
    """Interface of the reference's regulariser (utils/utils.py:293-312): what train.py:340 passes."""

    def __init__(self, TVLoss_weight=1):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x):
        tv = 0
        if x.size(2) > 1:
            tv = tv + torch.pow(x[:, :, 1:, :] - x[:, :, :-1, :], 2).mean()
        if x.size(3) > 1:
            tv = tv + torch.pow(x[:, :, :, 1:] - x[:, :, :, :-1], 2).mean()
        return self.TVLoss_weight * 2 * tv


@pytest.fixture()
def sched():
    g = load_golden("sched_nc")
    return g, module_from_golden(g)


def _scale_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", ["l1", "tv_density", "tv_app"])
def test_regularisers_value_and_grads(sched, name):
    g, m = sched
    fn = {"l1": m.density_L1, "tv_density": lambda: m.TV_loss_density(TVLoss()),
          "tv_app": lambda: m.TV_loss_app(TVLoss())}[name]
    m.zero_grad()
    torch.cuda.reset_peak_memory_stats()
    val = fn()
    val.backward()
    assert abs(float(val) - float(g[f"{name}.value"])) <= 2e-6 * abs(float(g[f"{name}.value"]))
    n_checked = 0
    for k, p in m.named_parameters():
        key = f"{name}.grad.{k}"
        if key in g:
            assert p.grad is not None, k
            assert _scale_err(p.grad.cpu().numpy(), g[key]) < 2e-5, k
            n_checked += 1
        else:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, k
    assert n_checked == 6


def test_tv_generic_callable_falls_back(sched):
    g, m = sched
    val = m.TV_loss_density(lambda x: TVLoss()(x))           # not a TVLoss instance: the reference's formulation
    assert abs(float(val) - float(g["tv_density.value"])) <= 2e-6 * abs(float(g["tv_density.value"]))


def test_density_l1_streams_at_640():
    """No G^3-sized allocation: at 640^3 the reference's bmm intermediate alone is 8.4 GB."""
    import bench
    lt = bench.build_scene("cuda", 640)
    rf = lt.tensorfs[0]
    torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    val = rf.density_L1()
    val.backward()
    torch.cuda.synchronize()
    extra = torch.cuda.max_memory_allocated() - base
    assert extra < 64 << 20, extra                              # gradient buffers of the density planes only
    assert torch.isfinite(val) and float(rf.density_plane[0].grad.abs().sum()) > 0


def test_alpha_mask_rebuild(sched):
    g, m = sched
    a0 = m.getDenseAlpha((10, 12, 14)).cpu().numpy()
    assert a0.shape == (10, 12, 14)
    # alpha = 1 - exp(-sigma * step) is quantised to the fp32 spacing below 1.0 (6e-8): two quanta
    assert np.abs(a0 - g["mask.alpha0"]).max() < 1.3e-7
    m.updateAlphaMask((10, 12, 14))
    vol = m.alphaMask.alpha_volume.cpu().numpy()
    ref = g["mask.volume"]
    assert vol.shape == ref.shape == (1, 1, 14, 12, 10)
    diff = np.argwhere(vol != ref)
    # a voxel may only differ where the pooled alpha sits within fp32 noise of alphaMask_thres
    pooled = g["mask.pooled"]
    for idx in diff:
        assert abs(pooled[tuple(idx[2:])] - 1e-4) < 1e-9, (idx, pooled[tuple(idx[2:])])
    assert len(diff) <= 2
    np.testing.assert_array_equal(m.alphaMask.aabb.cpu().numpy(), g["alphaMask.aabb"] if "alphaMask.aabb" in g
                                  else m.aabb.cpu().numpy())
    # with the mask in place the lattice evaluation culls like compute_alpha (tensorBase.py:538-558)
    if len(diff) == 0:
        a1 = m.getDenseAlpha((11, 9, 13)).cpu().numpy()
        assert np.abs(a1 - g["mask.alpha1"]).max() < 1.3e-7


def test_upsample_matches_reference(sched):
    g, m = sched
    m.upsample_volume_grid([30, 33, 41])
    assert m.nSamples == int(g["up.nSamples"]) and m._grid_host == [30, 33, 41]
    for k, v in m.state_dict().items():
        if "plane" in k or "line" in k:
            ref = g["up." + k]
            assert tuple(v.shape) == ref.shape, k
            assert np.abs(v.cpu().numpy() - ref).max() < 1e-6, k
    assert m.app_plane[1].is_contiguous(memory_format=torch.channels_last)
    # the upsampled field renders (pointers / layouts are what the kernel expects)
    rays = torch.from_numpy(g["sr.rays"]).cuda()
    with torch.no_grad():
        rgb, depth = m(rays)
    assert torch.isfinite(rgb).all() and torch.isfinite(depth).all()


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_sample_ray_aabb(sched, mode):
    g, m = sched
    rays = torch.from_numpy(g["sr.rays"]).cuda()
    if mode == "eval":
        pts, z, inside = m.sample_ray(rays[:, :3], rays[:, 3:], is_train=False, N_samples=40)
    else:
        torch.manual_seed(52)
        pts, z, inside = m.sample_ray(rays[:, :3], rays[:, 3:], is_train=True, N_samples=-1)
    ref_z, ref_pts, ref_in = g[f"sr.{mode}.z"], g[f"sr.{mode}.pts"], g[f"sr.{mode}.inside"]
    assert tuple(z.shape) == ref_z.shape
    np.testing.assert_allclose(z.cpu().numpy(), ref_z, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(pts.cpu().numpy(), ref_pts, rtol=1e-6, atol=2e-5)
    mism = inside.cpu().numpy() != ref_in
    # the mask may only differ where a point sits on a face of the box to rounding
    assert mism.sum() <= 2, int(mism.sum())
