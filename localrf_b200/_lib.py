"""ctypes binding of the C ABI (include/localrf_b200.h) + the in-tree nvcc build.

The product path has NO fallback: if the shared library is missing or cannot be loaded,
`lib()` raises.  Nothing here imports oracle/.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_CSRC, "liblrf_b200.so")
SOURCES = ["lrf_render.cu", "lrf_aux.cu", "lrf_grad.cu", "lrf_backward.cu", "lrf_sched.cu", "lrf_abi.cu"]
HEADERS = ["lrf_common.cuh", "lrf_device.cuh", "lrf_backward_tc.cuh", os.path.join("..", "..", "include", "localrf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

_fp = C.POINTER(C.c_float)
_vp = C.c_void_p


class LrfField(C.Structure):
    _fields_ = [
        ("grid", C.c_int32 * 3),
        ("aabb", C.c_float * 6),
        ("n_dcomp", C.c_int32), ("n_acomp", C.c_int32),
        ("dplane", _vp * 3), ("dline", _vp * 3), ("aplane", _vp * 3), ("aline", _vp * 3),
        ("app_dim", C.c_int32),
        ("basis", _vp),
        ("featureC", C.c_int32), ("fea_pe", C.c_int32), ("view_pe", C.c_int32),
        ("w1", _vp), ("b1", _vp), ("w2", _vp), ("b2", _vp), ("w3", _vp), ("b3", _vp),
        ("alpha_vol", _vp),
        ("alpha_dims", C.c_int32 * 3),
        ("alpha_aabb", C.c_float * 6),
        ("density_shift", C.c_float), ("distance_scale", C.c_float), ("weight_thres", C.c_float),
        ("act", C.c_int32),
        ("z_vals", _vp),
        ("n_samples", C.c_int32),
        ("grid_dtype", C.c_int32),       # GRID_F32 / GRID_BF16
    ]


GRID_F32, GRID_BF16 = 0, 1


class LrfBatch(C.Structure):
    _fields_ = [
        ("n_rays", C.c_int64),
        ("rays", _vp),
        ("ray_ids", _vp),
        ("W", C.c_int32), ("H", C.c_int32),
        ("fov360", C.c_int32),
        ("focal", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("intrinsics", _vp),
        ("cam2world", _vp),
        ("n_views", C.c_int64),
        ("world2rf", _vp),
        ("blend", _vp),
        ("blend_stride", C.c_int64),
        ("exposure", _vp),
        ("accumulate", C.c_int32), ("finalize", C.c_int32), ("white_bg", C.c_int32),
        ("floater_thresh", C.c_float), ("refine", C.c_int32),
    ]


class LrfOutputs(C.Structure):
    _fields_ = [("rgb", _vp), ("depth", _vp), ("weights", _vp), ("directions", _vp),
                ("ij", _vp), ("pix", _vp), ("stats", _vp),
                ("n_peers", C.c_int32), ("peer_pix", _vp * 16), ("mc_pix", _vp),
                ("peer_flags", _vp * 16), ("rank", C.c_int32), ("signal_seq", C.c_uint64), ("wait_seq", C.c_uint64)]


class LrfGradients(C.Structure):
    _fields_ = [("d_rays", _vp), ("d_dplane", _vp * 3), ("d_dline", _vp * 3), ("d_aplane", _vp * 3),
                ("d_aline", _vp * 3), ("d_w1b", _vp), ("d_b1", _vp), ("d_w2", _vp), ("d_b2", _vp),
                ("d_w3", _vp), ("d_b3", _vp)]


EXPORTS = ["lrf_version", "lrf_sizeof", "lrf_last_error", "lrf_prepared_bytes", "lrf_prepared_bytes_for", "lrf_field_prepare",
           "lrf_render", "lrf_mlp_forward", "lrf_app_products", "lrf_density_feature_backward",
           "lrf_app_products_backward", "lrf_density_feature", "lrf_app_feature", "lrf_repack_nchw_to_nhwc",
           "lrf_launch_info", "lrf_prepared_backward_bytes", "lrf_backward_scratch_bytes",
           "lrf_field_prepare_backward", "lrf_render_backward", "lrf_peer_barrier", "lrf_peer_signal_wait",
           "lrf_alpha_mask_build", "lrf_upsample", "lrf_density_l1", "lrf_density_l1_backward", "lrf_tv_sums",
           "lrf_tv_sums_backward", "lrf_sample_ray", "lrf_frame_to_u8", "lrf_pack_bf16"]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(_CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> localrf_b200/csrc/liblrf_b200.so"""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("LRF_NVCC_EXTRA", "").split()          # tuning experiments (e.g. -DLRF_THREADS=640)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)  # the image's CC points at a gcc without libgomp
    r = subprocess.run(cmd, cwd=_CSRC, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB_PATH


_lib = None


ABI_VERSION = 3


def lib():
    """Loads the shared library (building it if sources are newer).  Raises if it is missing, if a
    needed rebuild fails (a stale binary is never loaded silently), or if the binary's ABI version or
    struct sizes differ from the ctypes mirrors in this file."""
    global _lib
    if _lib is not None:
        return _lib
    if _stale():
        try:
            build()
        except (RuntimeError, FileNotFoundError) as e:
            raise RuntimeError(
                "localrf_b200: the CUDA library is missing or older than its sources and rebuilding "
                f"it failed; there is no CPU fallback and no stale binary is loaded ({e})") from e
    L = C.CDLL(LIB_PATH)
    if not hasattr(L, "lrf_sizeof") or L.lrf_version() != ABI_VERSION:
        raise RuntimeError(f"localrf_b200: {LIB_PATH} has a different ABI version than this binding; "
                           "rebuild it (python -c 'import localrf_b200; localrf_b200.build(force=True)')")
    L.lrf_sizeof.restype = C.c_size_t
    L.lrf_sizeof.argtypes = [C.c_int32]
    for which, mirror in enumerate((LrfField, LrfBatch, LrfOutputs, LrfGradients)):
        if L.lrf_sizeof(which) != C.sizeof(mirror):
            raise RuntimeError(f"localrf_b200: struct {mirror.__name__} is {L.lrf_sizeof(which)} bytes in "
                               f"{LIB_PATH} but {C.sizeof(mirror)} in the ctypes mirror -- stale build")
    L.lrf_version.restype = C.c_int
    L.lrf_last_error.restype = C.c_char_p
    L.lrf_prepared_bytes.restype = C.c_size_t
    L.lrf_prepared_bytes_for.restype = C.c_size_t
    L.lrf_prepared_bytes_for.argtypes = [C.POINTER(LrfField)]
    L.lrf_field_prepare.argtypes = [C.POINTER(LrfField), _vp, _vp]
    L.lrf_render.argtypes = [C.POINTER(LrfField), _vp, C.POINTER(LrfBatch), C.POINTER(LrfOutputs), _vp]
    L.lrf_mlp_forward.argtypes = [_vp, _vp, _vp, C.c_int64, _vp, _vp]
    L.lrf_density_feature.argtypes = [C.POINTER(LrfField), _vp, C.c_int64, _vp, _vp]
    L.lrf_app_feature.argtypes = [C.POINTER(LrfField), _vp, C.c_int64, _vp, _vp]
    L.lrf_app_products.argtypes = [C.POINTER(LrfField), _vp, C.c_int64, _vp, _vp]
    for name in ("lrf_density_feature_backward", "lrf_app_products_backward"):
        getattr(L, name).argtypes = [C.POINTER(LrfField), _vp, _vp, C.c_int64, _vp * 3, _vp * 3, _vp, _vp]
    L.lrf_repack_nchw_to_nhwc.argtypes = [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp]
    L.lrf_pack_bf16.argtypes = [_vp, _vp, C.c_int64, _vp]
    L.lrf_launch_info.argtypes = [C.POINTER(C.c_int32)] * 3
    L.lrf_prepared_backward_bytes.restype = C.c_size_t
    L.lrf_backward_scratch_bytes.restype = C.c_size_t
    L.lrf_backward_scratch_bytes.argtypes = [C.c_int64, C.c_int32]
    L.lrf_field_prepare_backward.argtypes = [C.POINTER(LrfField), _vp, _vp]
    L.lrf_render_backward.argtypes = [C.POINTER(LrfField), _vp, _vp, C.c_int64, C.c_int32, _vp, _vp,
                                      C.POINTER(LrfGradients), _vp, C.c_size_t, _vp]
    L.lrf_alpha_mask_build.argtypes = [C.POINTER(LrfField), C.c_int32 * 3, C.c_float, C.c_float, _vp, _vp, _vp, _vp]
    L.lrf_upsample.argtypes = [_vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32, _vp]
    L.lrf_density_l1.argtypes = [C.POINTER(LrfField), _vp, _vp]
    L.lrf_density_l1_backward.argtypes = [C.POINTER(LrfField), _vp, _vp * 3, _vp * 3, _vp]
    L.lrf_tv_sums.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp]
    L.lrf_tv_sums_backward.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, _vp, C.c_float, C.c_float, _vp, _vp]
    L.lrf_sample_ray.argtypes = [_vp, _vp, C.c_int64, C.c_int32, C.c_float * 6, C.c_float, C.c_float, C.c_float,
                                 _vp, _vp, _vp, _vp]
    L.lrf_frame_to_u8.argtypes = [_vp, C.c_int32, _vp, C.c_int32, C.c_int64, C.c_float, C.c_float, _vp, _vp, _vp, _vp]
    L.lrf_peer_barrier.argtypes = [C.POINTER(_vp), C.c_int32, C.c_int32, C.c_uint64, _vp]
    L.lrf_peer_signal_wait.argtypes = [C.POINTER(_vp), C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, _vp]
    for name in EXPORTS:
        getattr(L, name)  # AttributeError if the header and the library ever drift apart
    _lib = L
    return L


class LrfError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = lib().lrf_last_error().decode()
        if rc == -2:
            raise NotImplementedError(f"localrf_b200: {msg}")
        if rc == -1:
            raise ValueError(f"localrf_b200: {msg}")
        raise LrfError(f"localrf_b200 (code {rc}): {msg}")
