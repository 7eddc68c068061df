"""-m gpu: BASELINE-size parity (300^3 grid, 4096-ray batches, S = 344; 3-field blend; 640^3 with the
floater filter): EVERY ray of whole batches against the unmodified reference's committed outputs and
against the pinned CPU oracle, plus size-independent properties.  Rays that sit exactly on the
reference's hard w > 1e-3 switch are accounted for one by one (helpers.check_with_ties)."""
import numpy as np
import pytest
import torch

from gpu_helpers import oracle_fields, oracle_local
from helpers import check_with_ties, load_golden, rel_err

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


def _scene(n_fields, grid=300):
    """bench.build_scene + (config 3) two more fields with seeds 1, 2, as tests/golden/make_golden.py."""
    import bench
    lt = bench.build_scene("cpu", grid)
    for k in range(1, n_fields):
        lt.append_frame()                      # append_rf needs >= 2 frames to cross-fade over
        torch.manual_seed(k)
        lt.append_rf(1)
    return lt.to("cuda")


def _render(lt, ids, **kw):
    with torch.no_grad():
        rgb, depth, dirs, _ = lt(ids, torch.tensor([0], device="cuda"), 800, 800, is_train=False, **kw)
    return rgb.cpu().numpy(), depth.cpu().numpy(), dirs.cpu().numpy()


def test_cfg2_all_rays_vs_reference_golden(scene300):
    """Config 2, ALL 4096 rays of the first / a middle / the last full batch of the frame against the
    outputs of the unmodified reference (tests/golden/cfg2_300.npz, fields regenerated from the seed)."""
    g = load_golden("cfg2_300")
    for b in g["batches"]:
        rgb, depth, _ = _render(scene300, _frame_batch(int(b)))
        n1, e1 = check_with_ties(rgb, g[f"b{b}.rgb"], g[f"b{b}.margin"], TOL, f"cfg2 b{b} rgb")
        n2, e2 = check_with_ties(depth, g[f"b{b}.depth"], g[f"b{b}.margin"], TOL, f"cfg2 b{b} depth")
        print(f"cfg2 batch {b} vs reference: rgb worst {e1:.2e} ({n1} threshold ties), depth worst {e2:.2e}")
        assert n2 == 0


def test_cfg2_all_rays_vs_oracle(scene300):
    """Two more whole batches against the pinned CPU oracle (every ray, not a subset)."""
    fields = oracle_fields(scene300)
    for b in (40, 120):
        ids = _frame_batch(b)
        rgb, depth, dirs = _render(scene300, ids)
        ref = oracle_local(scene300, fields, ids.cpu().numpy(), 0, 800, 800)
        n1, e1 = check_with_ties(rgb, ref["rgb"], ref["margin"], TOL, f"cfg2 b{b} rgb")
        n2, e2 = check_with_ties(depth, ref["depth"], ref["margin"], TOL, f"cfg2 b{b} depth")
        print(f"cfg2 batch {b} vs oracle: rgb worst {e1:.2e} ({n1} threshold ties), depth worst {e2:.2e}")
        assert n2 == 0
        np.testing.assert_allclose(dirs, ref["directions"], rtol=2e-6, atol=1e-6)


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


def test_cfg3_three_fields_vs_reference_golden_and_oracle():
    """Config 3: three 300^3 fields (seeds 0,1,2), world2rf offsets, explicit blend [0.2,0.5,0.3]:
    all 4096 rays against the unmodified reference's outputs and against the oracle."""
    g = load_golden("cfg3_300")
    lt = _scene(3)
    w2rf = [torch.tensor(w, device="cuda") for w in g["world2rf"]]
    bw = torch.tensor(g["blend"], device="cuda")
    ids = _frame_batch(int(g["batch"]))
    rgb, depth, _ = _render(lt, ids, world2rf=w2rf, blending_weights=bw)
    # (a CPU run of the reference and a GPU run round exp() differently in the last bit often enough that
    #  ~0.5 % of the rays per field sit on a flipped w > 1e-3 decision; against the reference on the SAME
    #  GPU -- tests/test_gpu_vs_reference.py -- there are 0-1 such rays per batch)
    n1, e1 = check_with_ties(rgb, g["rgb"], g["margin"], TOL, "cfg3 rgb vs reference", max_tie_frac=0.03)
    n2, e2 = check_with_ties(depth, g["depth"], g["margin"], TOL, "cfg3 depth vs reference")
    print(f"cfg3 vs reference: rgb worst {e1:.2e} ({n1} threshold ties), depth worst {e2:.2e}")
    assert n2 == 0
    ids = _frame_batch(100)
    rgb, depth, _ = _render(lt, ids, world2rf=w2rf, blending_weights=bw)
    ref = oracle_local(lt, oracle_fields(lt), ids.cpu().numpy(), 0, 800, 800, world2rf=g["world2rf"],
                       blend=g["blend"])
    n1, e1 = check_with_ties(rgb, ref["rgb"], ref["margin"], TOL, "cfg3 rgb vs oracle", max_tie_frac=0.03)
    n2, e2 = check_with_ties(depth, ref["depth"], ref["margin"], TOL, "cfg3 depth vs oracle")
    print(f"cfg3 vs oracle: rgb worst {e1:.2e} ({n1} threshold ties), depth worst {e2:.2e}")
    assert n2 == 0


def test_cfg5_miniature_two_fields_640_floater():
    """Config 5 in miniature: two 640^3 fields (S = 738), natural cross-fade row, floater_thresh 0.5
    (train.py:107,139), 1024 rays against the oracle."""
    lt = _scene(2, grid=640)
    assert lt.tensorfs[0].sample_table(False, -1, torch.device("cuda")).numel() == 738
    w2rf = [torch.zeros(3, device="cuda"), torch.tensor([-0.3, 0.0, 0.1], device="cuda")]
    bw = torch.tensor([[0.4, 0.6]], device="cuda")
    ids = torch.arange(300 * 800 + 100, 300 * 800 + 100 + 1024, dtype=torch.int64, device="cuda")
    rgb, depth, _ = _render(lt, ids, world2rf=w2rf, blending_weights=bw, floater_thresh=0.5)
    ref = oracle_local(lt, oracle_fields(lt), ids.cpu().numpy(), 0, 800, 800,
                       world2rf=[w.cpu().numpy() for w in w2rf], blend=bw.cpu().numpy(), floater_thresh=0.5)
    n1, e1 = check_with_ties(rgb, ref["rgb"], ref["margin"], TOL, "cfg5-mini rgb", max_tie_frac=0.02)
    e2 = rel_err(depth, ref["depth"])
    print(f"cfg5 miniature vs oracle: rgb worst {e1:.2e} ({n1} threshold ties), depth {e2:.2e}")
    assert e2 < TOL
