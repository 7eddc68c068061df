"""CPU-only: the oracle's analytic backward (oracle/lrf_oracle.c::orc_field_backward) pinned against
the reference's own autograd (tests/golden/grads_field.npz, made by make_golden.py::gen_grads)."""
import numpy as np

from helpers import full_field_dict, load_golden
from oracle import oracle as orc


def scale_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def test_field_backward_vs_reference_autograd():
    g = load_golden("grads_field")
    fd = full_field_dict(g)
    f = orc.Field(fd)
    grads = orc.field_backward(f, fd, g["rays"], g["z"], g["c_rgb"], g["c_depth"], white_bg=True)
    assert scale_err(grads["rays"], g["grad.rays"]) < 1e-4
    checked = 0
    for key, val in grads.items():
        if key == "rays":
            continue
        ref = g["grad." + key]
        assert val.shape == ref.shape, key
        assert scale_err(val, ref) < 1e-4, (key, scale_err(val, ref))
        checked += 1
    assert checked == 19
