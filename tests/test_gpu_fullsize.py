"""-m gpu: BASELINE-size checks (300^3 grid, 4096-ray batches, S = 344) through size-independent
properties plus an oracle spot check on a subset of the rays."""
import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def scene300():
    import bench
    lt = bench.build_scene(torch.device("cuda", 0), 300)
    return lt


def _frame_batch(k, n=4096):
    return torch.arange(k * n, (k + 1) * n, dtype=torch.int64, device="cuda")


def test_sample_count_and_stats(scene300):
    lt = scene300
    rf = lt.tensorfs[0]
    assert rf.nSamples == 1036 and rf.sample_table(False, -1, torch.device("cuda")).numel() == 344
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        lt(_frame_batch(80), torch.tensor([0], device="cuda"), 800, 800, is_train=False, stats=stats)
    marched, shaded = stats.tolist()
    assert marched == 4096 * 343                       # no early termination at random init
    assert 0.40 < shaded / (4096 * 344) < 0.50         # K/S = 0.446


def test_subset_vs_oracle(scene300):
    """96 rays of a 4096-ray batch against the pinned CPU oracle at full grid size."""
    import bench
    lt = scene300
    ids = _frame_batch(77)
    with torch.no_grad():
        rgb, depth, dirs, ij = lt(ids, torch.tensor([0], device="cuda"), 800, 800, is_train=False)
    sel = np.arange(0, 4096, 43)[:96]
    field = bench.oracle_field(lt)
    ref = bench.oracle_batch(lt, field, ids.cpu().numpy()[sel])
    assert rel_err(rgb.cpu().numpy()[sel], ref["rgb"]) < TOL
    assert rel_err(depth.cpu().numpy()[sel], ref["depth"]) < TOL
    np.testing.assert_allclose(dirs.cpu().numpy()[sel], ref["directions"], rtol=2e-6, atol=1e-6)


def test_batch_split_and_order_bit_exact(scene300):
    lt = scene300
    v = torch.tensor([0], device="cuda")
    ids = _frame_batch(40)
    with torch.no_grad():
        rgb, depth, _, _ = lt(ids, v, 800, 800, is_train=False)
        r1, d1, _, _ = lt(ids[:1000], v, 800, 800, is_train=False)
        r2, d2, _, _ = lt(ids[1000:], v, 800, 800, is_train=False)
        perm = torch.randperm(4096, generator=torch.Generator().manual_seed(3)).cuda()
        rp, dp, _, _ = lt(ids[perm], v, 800, 800, is_train=False)
    assert torch.equal(torch.cat([r1, r2]), rgb) and torch.equal(torch.cat([d1, d2]), depth)
    assert torch.equal(rp, rgb[perm]) and torch.equal(dp, depth[perm])


def test_weights_are_a_partition_of_unity(scene300):
    """alpha[:, -1] = 1 makes every ray's weights sum to 1 (acc ~= 1, tensorBase.py:24)."""
    rf = scene300.tensorfs[0]
    g = torch.Generator().manual_seed(9)
    rays = torch.cat([0.1 * torch.randn(512, 3, generator=g), torch.randn(512, 3, generator=g)], -1).cuda()
    with torch.no_grad():
        rf(rays, return_weights=True, floater_thresh=0.5)
    s = rf.last_weights.sum(-1)
    assert float((s - 1).abs().max()) < 1e-5


def test_three_field_blend_is_linear():
    """Config 3: blend [0.2,0.5,0.3] of three 300^3 fields == the weighted sum of the single-field
    renders (exposure identity, no clamping active at random init)."""
    import bench
    import localrf_b200 as L
    torch.manual_seed(0)
    lt = bench.build_scene("cpu", 300)
    for k in (1, 2):
        lt.append_frame()                      # append_rf needs >= 2 frames to cross-fade over
        torch.manual_seed(k)
        lt.append_rf(1)
    lt = lt.to("cuda")
    w2rf = [torch.zeros(3, device="cuda"), torch.tensor([-0.3, 0.0, 0.0], device="cuda"),
            torch.tensor([-0.6, 0.0, 0.0], device="cuda")]
    ids = _frame_batch(60, 2048)
    v = torch.tensor([0], device="cuda")
    with torch.no_grad():
        def run(bw):
            return lt(ids, v, 800, 800, is_train=False, world2rf=w2rf,
                      blending_weights=torch.tensor([bw], device="cuda"))[:2]
        rgb, depth = run([0.2, 0.5, 0.3])
        parts = [run([1.0 if j == k else 0.0 for j in range(3)]) for k in range(3)]
    exp_rgb = sum(w * p[0] for w, p in zip([0.2, 0.5, 0.3], parts))
    exp_depth = sum(w * p[1] for w, p in zip([0.2, 0.5, 0.3], parts))
    assert rel_err(rgb.cpu().numpy(), exp_rgb.cpu().numpy()) < 1e-5
    assert rel_err(depth.cpu().numpy(), exp_depth.cpu().numpy()) < 1e-5
    assert float((parts[0][0] - parts[1][0]).abs().max()) > 1e-3   # the fields really differ
