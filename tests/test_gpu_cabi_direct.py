"""-m gpu: the raw C ABI, driven exactly like the ctypes stub of INTEGRATION.md §2 (hand-filled
LrfField / LrfBatch / LrfOutputs, no Python mirror in between), against the pinned oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import full_field_dict, load_golden, rel_err
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_render_rays_through_raw_abi():
    from localrf_b200._lib import LIB_PATH, LrfBatch, LrfField, LrfOutputs
    L = C.CDLL(LIB_PATH)
    L.lrf_last_error.restype = C.c_char_p
    L.lrf_prepared_bytes.restype = C.c_size_t
    g = load_golden("opaque_32")
    dev = torch.device("cuda")
    keep = []

    def chan_last(key):                       # reference [1,C,H,W] -> device [H][W][C]
        t = torch.from_numpy(g[key]).to(dev).permute(0, 2, 3, 1).contiguous(); keep.append(t); return t.data_ptr()

    def plain(key):
        t = torch.from_numpy(g[key]).to(dev).contiguous(); keep.append(t); return t.data_ptr()

    f = LrfField()
    f.grid[:] = [int(v) for v in g["gridSize"]]
    f.aabb[:] = [float(v) for v in g["aabb"].reshape(-1)]
    f.n_dcomp, f.n_acomp = 8, 24
    for i in range(3):
        f.dplane[i] = chan_last(f"density_plane.{i}"); f.dline[i] = chan_last(f"density_line.{i}")
        f.aplane[i] = chan_last(f"app_plane.{i}"); f.aline[i] = chan_last(f"app_line.{i}")
    f.app_dim, f.featureC, f.fea_pe, f.view_pe = 27, 128, 0, 0
    f.basis = plain("basis_mat.weight")
    f.w1, f.b1 = plain("renderModule.mlp.0.weight"), plain("renderModule.mlp.0.bias")
    f.w2, f.b2 = plain("renderModule.mlp.2.weight"), plain("renderModule.mlp.2.bias")
    f.w3, f.b3 = plain("renderModule.mlp_view.0.weight"), plain("renderModule.mlp_view.0.bias")
    f.density_shift, f.distance_scale, f.weight_thres, f.act = float(g["density_shift"]), 25.0, 1e-3, 0
    z = torch.from_numpy(g["eval.z"]).to(dev)
    f.z_vals, f.n_samples = z.data_ptr(), z.numel()
    prepared = torch.empty(L.lrf_prepared_bytes(), dtype=torch.uint8, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.lrf_field_prepare(C.byref(f), C.c_void_p(prepared.data_ptr()), stream) == 0, L.lrf_last_error()
    rays = torch.from_numpy(g["rays"]).to(dev)
    n = rays.shape[0]
    rgb = torch.empty(n, 3, device=dev); depth = torch.empty(n, device=dev)
    w = torch.empty(n, z.numel(), device=dev)
    b = LrfBatch(); b.n_rays = n; b.rays = rays.data_ptr(); b.n_views = 1; b.white_bg = 1
    o = LrfOutputs(); o.rgb, o.depth, o.weights = rgb.data_ptr(), depth.data_ptr(), w.data_ptr()
    assert L.lrf_render(C.byref(f), C.c_void_p(prepared.data_ptr()), C.byref(b), C.byref(o), stream) == 0, L.lrf_last_error()
    torch.cuda.synchronize()
    assert rel_err(rgb.cpu().numpy(), g["eval.rgb"]) < 1e-4
    assert rel_err(depth.cpu().numpy(), g["eval.depth"]) < 1e-4
    ref = orc.field_forward(orc.Field(full_field_dict(g)), g["rays"], g["eval.z"])
    assert rel_err(w.cpu().numpy(), ref["weights"], floor=5e-3) < 1e-4
    # error convention: NULL prepared block -> LRF_ERR_INVALID with a message
    assert L.lrf_render(C.byref(f), None, C.byref(b), C.byref(o), stream) == -1
    assert b"prepared" in L.lrf_last_error()
