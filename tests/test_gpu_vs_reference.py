"""-m gpu: the fused path against the UNMODIFIED reference running on the SAME GPU, same parameters,
same rays -- whole 4096-ray batches at the BASELINE size, every ray compared.

The reference files travel to the GPU box as a byte-identical staged copy (baseline/_ref, written by
oracle/vendor_ref.py in the build container; git-ignored).  Skipped when that copy is absent.
Both sides use CUDA's expf here, so -- unlike comparisons with a CPU run -- alpha = 1 - exp(-x) is
computed from the same exp; remaining differences are summation order (grid_sample / cumprod / GEMM)."""
import contextlib
import io

import numpy as np
import pytest
import torch

import bench
from helpers import check_with_ties, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ref_mod():
    from oracle.ref_loader import reference_available
    if not reference_available():
        pytest.skip("no staged reference (baseline/_ref) on this box")
    cls = bench.load_reference_classes()
    import models.tensorBase as ref_tb          # the reference's module (oracle/ref_loader put it on sys.path)
    cap = {}
    orig = ref_tb.alpha2weights

    def a2w(alpha):                               # wrap, don't modify: capture the weights it returns
        w, T = orig(alpha)
        cap.setdefault("w", []).append(w.detach())
        return w, T

    ref_tb.alpha2weights = a2w
    yield cls, cap
    ref_tb.alpha2weights = orig


def _pair(wl_name, ref_cls):
    import localrf_b200 as L
    wl = bench.Workload(wl_name)
    ref = bench.ReferenceRunner(wl, "cuda")
    ours = wl.build(L.LocalTensorfs, quiet=True).to("cuda")
    for a, b in zip(ref.lt.state_dict().values(), ours.state_dict().values()):
        assert torch.equal(a.cpu(), b.cpu())                 # same seed -> same init, both constructors
    return wl, ref, ours


@pytest.mark.parametrize("wl_name,batches", [("cfg2", (3, 90)), ("distB", (50,)), ("incoherent", (0, 7)),
                                             ("cfg3", (60,))])
def test_whole_batches_vs_reference_on_gpu(ref_mod, wl_name, batches):
    cls, cap = ref_mod
    wl, ref, ours = _pair(wl_name, cls)
    kw = wl.call_kwargs(ours, torch.device("cuda"))
    for b in batches:
        ids, view = ref.ids[b], ref.views[b]
        cap.clear()
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            r_rgb, r_depth, r_dirs, r_ij = ref.lt(ids, view, 800, 800, **ref.kw)
            rgb, depth, dirs, ij = ours(ids, view, 800, 800, **kw)
        # per-ray threshold margin from the reference's own weights (all chunks x fields)
        n_f = 3 if wl_name == "cfg3" else 1          # alpha2weights calls arrive chunk-major, field-minor
        per = [torch.cat(cap["w"][k::n_f]) for k in range(n_f)]
        margin = torch.stack([(w - 1e-3).abs().min(-1).values for w in per]).min(0).values.cpu().numpy()
        n1, e1 = check_with_ties(rgb.cpu().numpy(), r_rgb.cpu().numpy(), margin, TOL, f"{wl_name} b{b} rgb")
        n2, e2 = check_with_ties(depth.cpu().numpy(), r_depth.cpu().numpy(), margin, TOL, f"{wl_name} b{b} depth")
        print(f"{wl_name} batch {b} vs reference-on-GPU: rgb worst {e1:.2e} ({n1} threshold ties of "
              f"{ids.shape[0]} rays), depth worst {e2:.2e}")
        assert n2 == 0
        assert torch.equal(ij, r_ij)
        np.testing.assert_allclose(dirs.cpu().numpy(), r_dirs.cpu().numpy(), rtol=2e-6, atol=1e-6)


def test_weights_vs_reference_on_gpu(ref_mod):
    """Per-sample weights [4096, 344] of one field against the reference's, same GPU."""
    cls, cap = ref_mod
    wl, ref, ours = _pair("cfg2", cls)
    g = torch.Generator().manual_seed(4)
    rays = torch.cat([0.1 * torch.randn(4096, 3, generator=g), torch.randn(4096, 3, generator=g)], -1).cuda()
    cap.clear()
    with torch.no_grad():
        r_rgb, r_depth = ref.lt.tensorfs[0](rays, is_train=False, white_bg=True, N_samples=-1)
        rgb, depth = ours.tensorfs[0](rays, is_train=False, return_weights=True)
    w_ref, w = cap["w"][-1].cpu().numpy(), ours.tensorfs[0].last_weights.cpu().numpy()
    e_floor = rel_err(w, w_ref, floor=5e-3)
    e_pure = rel_err(w, w_ref, floor=1e-3)
    print(f"weights vs reference-on-GPU: rel err {e_floor:.2e} (floor 5e-3), {e_pure:.2e} (floor 1e-3)")
    assert e_floor < TOL
    margin = np.abs(w_ref - 1e-3).min(-1)
    check_with_ties(rgb.cpu().numpy(), r_rgb.cpu().numpy(), margin, TOL, "random rays rgb")
    assert rel_err(depth.cpu().numpy(), r_depth.cpu().numpy()) < TOL
