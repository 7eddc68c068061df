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
