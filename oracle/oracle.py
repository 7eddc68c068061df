"""ctypes front-end of the CPU oracle (oracle/lrf_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under localrf_b200/ may import this module.

A "field" is a plain dict of numpy fp32 arrays in the REFERENCE layout, keyed like the reference's
``TensorVMSplit.state_dict()``, plus the scalar kwargs of its constructor:

    density_plane.{0,1,2} [1,C,H,W]   density_line.{0,1,2} [1,C,L,1]
    app_plane.{0,1,2}     [1,C,H,W]   app_line.{0,1,2}     [1,C,L,1]
    basis_mat.weight, renderModule.mlp.{0,2}.{weight,bias}, renderModule.mlp_view.0.{weight,bias}
    aabb [2,3]; optional alphaMask.alpha_volume [1,1,D,H,W], alphaMask.aabb [2,3]
    kwargs: gridSize, density_shift, distance_scale, rayMarch_weight_thres, fea2denseAct,
            fea_pe, view_pe, featureC, app_dim, step_ratio
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblrf_oracle.so")
_lib = None

_fp = C.POINTER(C.c_float)


class OrcField(C.Structure):
    _fields_ = [
        ("grid", C.c_int32 * 3),
        ("aabb", C.c_float * 6),
        ("n_dcomp", C.c_int32 * 3),
        ("n_acomp", C.c_int32 * 3),
        ("dplane", _fp * 3), ("dline", _fp * 3), ("aplane", _fp * 3), ("aline", _fp * 3),
        ("app_dim", C.c_int32),
        ("basis", _fp),
        ("featureC", C.c_int32), ("fea_pe", C.c_int32), ("view_pe", C.c_int32),
        ("w1", _fp), ("b1", _fp), ("w2", _fp), ("b2", _fp), ("w3", _fp), ("b3", _fp),
        ("alpha_vol", _fp),
        ("alpha_dims", C.c_int32 * 3),
        ("alpha_aabb", C.c_float * 6),
        ("density_shift", C.c_float), ("distance_scale", C.c_float), ("weight_thres", C.c_float),
        ("act", C.c_int32),
    ]


class OrcGrads(C.Structure):
    _fields_ = [("dplane", _fp * 3), ("dline", _fp * 3), ("aplane", _fp * 3), ("aline", _fp * 3),
                ("basis", _fp), ("w1", _fp), ("b1", _fp), ("w2", _fp), ("b2", _fp), ("w3", _fp),
                ("b3", _fp)]


def build(force=False):
    """Compile oracle/lrf_oracle.c -> oracle/liblrf_oracle.so (recipe == oracle/Makefile)."""
    src = os.path.join(_HERE, "lrf_oracle.c")
    hdr = os.path.join(_HERE, "lrf_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    base = ["-O2", "-fPIC", "-ffp-contract=off", "-fno-math-errno", "-std=c11", "-shared",
            "-o", _LIB_PATH, src, "-lm"]
    errs = []
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, capture_output=True, text=True)
                return _LIB_PATH
            except (subprocess.CalledProcessError, FileNotFoundError) as e:  # pragma: no cover
                errs.append(f"{cc} {omp}: {getattr(e, 'stderr', e)}")
    raise RuntimeError("could not build the oracle:\n" + "\n".join(errs))


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_sample_table.restype = C.c_int32
        L.orc_sample_table.argtypes = [C.c_int32, _fp, _fp, _fp]
        L.orc_contract.argtypes = [_fp, C.c_int64]
        for name in ("orc_density_feature", "orc_app_feature", "orc_alpha_mask_sample"):
            getattr(L, name).argtypes = [C.POINTER(OrcField), _fp, C.c_int64, _fp]
        L.orc_mlp_late_view.argtypes = [C.POINTER(OrcField), _fp, _fp, C.c_int64, C.c_int, _fp]
        L.orc_field_forward.argtypes = [
            C.POINTER(OrcField), _fp, C.c_int64, _fp, C.c_int32, C.c_int, C.c_float, C.c_int,
            _fp, _fp, _fp, _fp, C.POINTER(C.c_int32), C.c_int]
        L.orc_sixD_to_mtx.argtypes = [_fp, C.c_int64, _fp]
        L.orc_ray_directions.argtypes = [
            C.POINTER(C.c_int64), C.c_int64, C.c_int32, C.c_int32, C.c_int, C.c_float, C.c_float,
            C.c_float, _fp, C.POINTER(C.c_int64)]
        L.orc_local_forward.argtypes = [
            C.POINTER(OrcField), C.c_int32, C.POINTER(_fp), C.POINTER(C.c_int32),
            C.POINTER(C.c_int64), C.c_int64, C.c_int32, C.c_int32, C.c_int, C.c_float, C.c_float,
            C.c_float, _fp, C.c_int64, _fp, _fp, _fp, C.c_int, C.c_float, C.c_int, _fp, _fp, _fp,
            C.c_int]
        L.orc_local_forward_m.argtypes = L.orc_local_forward.argtypes[:-1] + [_fp, C.c_int]
        L.orc_field_backward.argtypes = [
            C.POINTER(OrcField), _fp, C.c_int64, _fp, C.c_int32, C.c_int, _fp, _fp,
            C.POINTER(OrcGrads), _fp]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return a.ctypes.data_as(_fp)


class Field:
    """Owns the contiguous fp32 copies the C struct points at."""

    def __init__(self, fd):
        self.keep = []
        s = OrcField()
        planes = [_f32(fd[f"density_plane.{i}"]) for i in range(3)]
        g = [0, 0, 0]
        # plane 0 is [1,C,G[1],G[0]], plane 1 is [1,C,G[2],G[0]]   (tensoRF.py:33-39)
        g[0], g[1], g[2] = planes[0].shape[3], planes[0].shape[2], planes[1].shape[2]
        if "gridSize" in fd:
            assert [int(v) for v in fd["gridSize"]] == g, (fd["gridSize"], g)
        aabb = _f32(fd["aabb"]).reshape(6)
        for i in range(3):
            s.grid[i] = g[i]
        for i in range(6):
            s.aabb[i] = float(aabb[i])
        for i in range(3):
            for name, arr_key, comp in (("dplane", f"density_plane.{i}", s.n_dcomp),
                                        ("dline", f"density_line.{i}", s.n_dcomp),
                                        ("aplane", f"app_plane.{i}", s.n_acomp),
                                        ("aline", f"app_line.{i}", s.n_acomp)):
                a = _f32(fd[arr_key])
                a = a.reshape(a.shape[1], -1)  # [C, H*W] / [C, L]
                self.keep.append(a)
                getattr(s, name)[i] = _p(a)
                comp[i] = a.shape[0]
        def keep(key):
            a = _f32(fd[key]); self.keep.append(a); return _p(a)
        s.basis = keep("basis_mat.weight")
        s.app_dim = int(np.asarray(fd["basis_mat.weight"]).shape[0])
        s.w1, s.b1 = keep("renderModule.mlp.0.weight"), keep("renderModule.mlp.0.bias")
        s.w2, s.b2 = keep("renderModule.mlp.2.weight"), keep("renderModule.mlp.2.bias")
        s.w3, s.b3 = keep("renderModule.mlp_view.0.weight"), keep("renderModule.mlp_view.0.bias")
        s.featureC = int(np.asarray(fd["renderModule.mlp.2.weight"]).shape[0])
        s.fea_pe = int(fd.get("fea_pe", 0))
        s.view_pe = int(fd.get("view_pe", 0))
        assert np.asarray(fd["renderModule.mlp.0.weight"]).shape[1] == s.app_dim * (1 + 2 * s.fea_pe)
        assert np.asarray(fd["renderModule.mlp_view.0.weight"]).shape[1] == s.featureC + 3 * (1 + 2 * s.view_pe)
        if fd.get("alphaMask.alpha_volume") is not None:
            v = _f32(fd["alphaMask.alpha_volume"])
            v = v.reshape(v.shape[-3:])
            self.keep.append(v)
            s.alpha_vol = _p(v)
            for i in range(3):
                s.alpha_dims[i] = v.shape[i]
            ab = _f32(fd["alphaMask.aabb"]).reshape(6)
            for i in range(6):
                s.alpha_aabb[i] = float(ab[i])
        s.density_shift = float(fd.get("density_shift", -10.0))
        s.distance_scale = float(fd.get("distance_scale", 25.0))
        s.weight_thres = float(fd.get("rayMarch_weight_thres", 1e-3))
        s.act = {"softplus": 0, "relu": 1}[str(fd.get("fea2denseAct", "softplus"))]
        self.struct = s
        self.grid = g
        self.step_ratio = float(fd.get("step_ratio", 0.5))

    # models/tensorBase.py:317-328  update_stepSize()
    def n_samples(self):
        aabb = np.array(list(self.struct.aabb), dtype=np.float32).reshape(2, 3)
        size = aabb[1] - aabb[0]
        units = size / (np.array(self.grid, dtype=np.float32) - 1)
        step = np.float32(np.mean(units, dtype=np.float32) * np.float32(self.step_ratio))
        diag = np.sqrt(np.sum(np.square(size), dtype=np.float32), dtype=np.float32)
        return int(np.float32(diag / step)) + 1


def sample_table(nSamples, j1=None, j2=None):
    N = nSamples // 6
    z = np.empty(2 * N, dtype=np.float32)
    a1 = _f32(j1).reshape(-1) if j1 is not None else None
    a2 = _f32(j2).reshape(-1) if j2 is not None else None
    lib().orc_sample_table(nSamples, _p(a1) if a1 is not None else None,
                           _p(a2) if a2 is not None else None, _p(z))
    return z


def contract(xyz):
    a = _f32(xyz).copy()
    lib().orc_contract(_p(a), a.size // 3)
    return a


def density_feature(field, xyz_norm):
    x = _f32(xyz_norm).reshape(-1, 3)
    out = np.empty(x.shape[0], dtype=np.float32)
    lib().orc_density_feature(C.byref(field.struct), _p(x), x.shape[0], _p(out))
    return out


def app_feature(field, xyz_norm):
    x = _f32(xyz_norm).reshape(-1, 3)
    out = np.empty((x.shape[0], field.struct.app_dim), dtype=np.float32)
    lib().orc_app_feature(C.byref(field.struct), _p(x), x.shape[0], _p(out))
    return out


def alpha_mask_sample(field, xyz):
    x = _f32(xyz).reshape(-1, 3)
    out = np.empty(x.shape[0], dtype=np.float32)
    lib().orc_alpha_mask_sample(C.byref(field.struct), _p(x), x.shape[0], _p(out))
    return out


def mlp_late_view(field, feat, viewdirs, refine=True):
    f = _f32(feat).reshape(-1, field.struct.app_dim)
    v = _f32(viewdirs).reshape(-1, 3)
    out = np.empty((f.shape[0], 3), dtype=np.float32)
    lib().orc_mlp_late_view(C.byref(field.struct), _p(f), _p(v), f.shape[0], int(refine), _p(out))
    return out


def field_forward(field, rays, z, white_bg=True, floater_thresh=0.0, refine=True, n_threads=0):
    """-> dict(rgb[N,3], depth[N], weights[N,S], acc[N], n_app[N])"""
    r = _f32(rays).reshape(-1, 6)
    z = _f32(z).reshape(-1)
    N, S = r.shape[0], z.shape[0]
    rgb = np.empty((N, 3), np.float32); depth = np.empty(N, np.float32)
    w = np.empty((N, S), np.float32); acc = np.empty(N, np.float32)
    n_app = np.empty(N, np.int32)
    lib().orc_field_forward(C.byref(field.struct), _p(r), N, _p(z), S, int(bool(white_bg)),
                            float(floater_thresh), int(bool(refine)), _p(rgb), _p(depth), _p(w),
                            _p(acc), n_app.ctypes.data_as(C.POINTER(C.c_int32)), int(n_threads))
    return dict(rgb=rgb, depth=depth, weights=w, acc=acc, n_app=n_app)


def sixD_to_mtx(r6):
    r = _f32(r6).reshape(-1, 3, 2)
    out = np.empty((r.shape[0], 3, 3), np.float32)
    lib().orc_sixD_to_mtx(_p(r), r.shape[0], _p(out))
    return out


def ray_directions(ray_ids, W, H, fov360, focal, cx, cy):
    ids = np.ascontiguousarray(ray_ids, dtype=np.int64)
    dirs = np.empty((ids.shape[0], 3), np.float32)
    ij = np.empty((ids.shape[0], 2), np.int64)
    lib().orc_ray_directions(ids.ctypes.data_as(C.POINTER(C.c_int64)), ids.shape[0], W, H,
                             int(fov360), float(focal), float(cx), float(cy), _p(dirs),
                             ij.ctypes.data_as(C.POINTER(C.c_int64)))
    return dirs, ij


def local_forward(fields, zs, ray_ids, W, H, fov360, focal, cx, cy, cam2world, world2rf, blend,
                  exposure=None, white_bg=True, floater_thresh=0.0, refine=False, n_threads=0,
                  with_margin=False):
    """LocalTensorfs.forward restated -> dict(rgb[N,3], depth[N], directions[N,3]); with_margin adds
    margin[N] = each ray's smallest |w - rayMarch_weight_thres| over samples and active fields."""
    n = len(fields)
    arr = (OrcField * n)(*[f.struct for f in fields])
    zs = [_f32(z).reshape(-1) for z in zs]
    zp = (_fp * n)(*[_p(z) for z in zs])
    Ss = (C.c_int32 * n)(*[z.shape[0] for z in zs])
    ids = np.ascontiguousarray(ray_ids, dtype=np.int64)
    N = ids.shape[0]
    c2w = _f32(cam2world).reshape(-1, 3, 4)
    V = c2w.shape[0]
    w2r = _f32(world2rf).reshape(n, 3)
    bl = _f32(blend).reshape(V, n)
    ex = _f32(exposure).reshape(V, 3, 3) if exposure is not None else None
    rgb = np.empty((N, 3), np.float32); depth = np.empty(N, np.float32)
    dirs = np.empty((N, 3), np.float32)
    margin = np.empty(N, np.float32) if with_margin else None
    lib().orc_local_forward_m(arr, n, zp, Ss, ids.ctypes.data_as(C.POINTER(C.c_int64)), N, W, H,
                              int(fov360), float(focal), float(cx), float(cy), _p(c2w), V, _p(w2r),
                              _p(bl), _p(ex) if ex is not None else None, int(bool(white_bg)),
                              float(floater_thresh), int(bool(refine)), _p(rgb), _p(depth), _p(dirs),
                              _p(margin) if with_margin else None, int(n_threads))
    out = dict(rgb=rgb, depth=depth, directions=dirs)
    if with_margin:
        out["margin"] = margin
    return out


def field_backward(field, fd, rays, z, g_rgb, g_depth, white_bg=True):
    """Analytic gradients of TensorBase.forward (floater 0) -> dict keyed like the reference's
    named_parameters (reference layouts) plus "rays".  `fd` is the field dict `field` was built from
    (for the parameter shapes)."""
    r = _f32(rays).reshape(-1, 6); z = _f32(z).reshape(-1)
    gr = _f32(g_rgb).reshape(-1, 3); gd = _f32(g_depth).reshape(-1)
    out = {}
    G = OrcGrads()

    def buf(key):
        a = np.zeros(np.asarray(fd[key]).shape, np.float32); out[key] = a; return _p(a)

    for i in range(3):
        G.dplane[i] = buf(f"density_plane.{i}"); G.dline[i] = buf(f"density_line.{i}")
        G.aplane[i] = buf(f"app_plane.{i}"); G.aline[i] = buf(f"app_line.{i}")
    G.basis = buf("basis_mat.weight")
    G.w1, G.b1 = buf("renderModule.mlp.0.weight"), buf("renderModule.mlp.0.bias")
    G.w2, G.b2 = buf("renderModule.mlp.2.weight"), buf("renderModule.mlp.2.bias")
    G.w3, G.b3 = buf("renderModule.mlp_view.0.weight"), buf("renderModule.mlp_view.0.bias")
    d_rays = np.zeros_like(r)
    lib().orc_field_backward(C.byref(field.struct), _p(r), r.shape[0], _p(z), z.shape[0],
                             int(bool(white_bg)), _p(gr), _p(gd), C.byref(G), _p(d_rays))
    out["rays"] = d_rays
    return out
