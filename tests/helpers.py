"""Shared test helpers: golden-fixture loading and error metrics.  (tests/ may use oracle/.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

FIELD_KEYS = (
    ["aabb"]
    + [f"{n}.{i}" for n in ("density_plane", "density_line", "app_plane", "app_line") for i in range(3)]
    + ["basis_mat.weight", "renderModule.mlp.0.weight", "renderModule.mlp.0.bias",
       "renderModule.mlp.2.weight", "renderModule.mlp.2.bias",
       "renderModule.mlp_view.0.weight", "renderModule.mlp_view.0.bias"]
)
SCALAR_KEYS = ("density_shift", "distance_scale", "rayMarch_weight_thres", "fea_pe", "view_pe",
               "featureC", "app_dim", "step_ratio")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def field_dict(g, prefix=""):
    """Extract one field's parameter dict (reference layout) from a golden file."""
    fd = {k: g[prefix + k] for k in FIELD_KEYS}
    for k in ("alphaMask.alpha_volume", "alphaMask.aabb"):
        if prefix + k in g:
            fd[k] = g[prefix + k]
    return fd


def field_scalars(g):
    sc = {k: g[k].item() for k in SCALAR_KEYS if k in g}
    sc["fea2denseAct"] = str(g["fea2denseAct"])
    sc["gridSize"] = [int(v) for v in g["gridSize"]]
    return sc


def full_field_dict(g, prefix=""):
    fd = field_dict(g, prefix)
    fd.update(field_scalars(g))
    return fd


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor): the parity metric (1e-4 relative, fp32)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


# ---- BASELINE-size parity with the reference's hard switch accounted for ------------------------
# tensorBase.py:622 shades a sample iff w > rayMarch_weight_thres (1e-3).  alpha = 1 - exp(-x) is
# quantised to the fp32 spacing below 1.0 (6e-8), so two correct fp32 evaluations of the same ray
# (ATen's vectorised exp, glibc's, CUDA's -- each within 1 ulp, not bit-identical) can put a weight
# that is within ~1e-7 of 1e-3 on different sides of the switch; the ray's colour then moves by that
# sample's w * rgb <= 1e-3.  Such a ray is an exact-threshold TIE, not a parity error: it is accepted
# only if its smallest |w - thres| (computed by the comparison's reference side) is below TIE_MARGIN
# and its error is bounded by a couple of flipped samples.
TIE_MARGIN = 2.5e-7
TIE_MAX_ERR = 2.5e-3


def ray_err(a, b, floor=1e-3):
    a = np.asarray(a, np.float64).reshape(len(a), -1)
    b = np.asarray(b, np.float64).reshape(len(b), -1)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor), axis=1)


def check_with_ties(out, ref, margin, tol=1e-4, what="", max_tie_frac=0.01):
    """Every ray within `tol` of the reference, except proven threshold ties.  Returns
    (n_ties, worst error among the non-tie rays)."""
    err = ray_err(out, ref)
    bad = np.nonzero(err > tol)[0]
    margin = np.asarray(margin, np.float64)
    unexplained = [(int(r), float(err[r]), float(margin[r])) for r in bad
                   if not (margin[r] < TIE_MARGIN and err[r] < TIE_MAX_ERR)]
    assert not unexplained, (f"{what}: {len(unexplained)} rays beyond {tol} that are NOT threshold ties "
                             f"(ray, err, min|w-thres|): {unexplained[:8]}")
    assert len(bad) <= max_tie_frac * len(err), f"{what}: {len(bad)} threshold ties of {len(err)} rays"
    ok = np.ones(len(err), bool); ok[bad] = False
    return len(bad), float(err[ok].max()) if ok.any() else 0.0
