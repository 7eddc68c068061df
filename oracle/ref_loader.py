"""Import the UNMODIFIED reference (facebookresearch/localrf) from /root/reference on CPU.

TEST / BENCH INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (in the build container, where
/root/reference is mounted) to produce the committed golden fixtures, and -- through the staged
byte-identical copy under baseline/_ref/ (oracle/vendor_ref.py) -- by `bench.py --impl reference` and
tests/test_gpu_vs_reference.py on the GPU box.  Nothing under localrf_b200/ imports this file.

The hot path's modules import a few packages that are missing from this image but are never
*used* on the path (kornia.create_meshgrid at utils/ray_utils.py:6, matplotlib / plyfile /
skimage.measure at utils/utils.py:10-12).  Empty stub modules are registered for those names so the
reference files import unmodified.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """$LOCALRF_REFERENCE, else /root/reference (build container), else the byte-identical staged
    copy under baseline/_ref/ (oracle/vendor_ref.py; the only one present on the GPU box)."""
    cands = [os.environ.get("LOCALRF_REFERENCE"), "/root/reference",
             os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "localTensoRF", "local_tensorfs.py")):
            return c
    return cands[1]


REF_ROOT = _find_root()
REF_PKG = os.path.join(REF_ROOT, "localTensoRF")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_PKG, "local_tensorfs.py"))


def _stub(name, **attrs):
    try:
        importlib.import_module(name)
        return
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so "import a.b" works
    sys.modules[name] = m
    if "." in name:
        parent, child = name.rsplit(".", 1)
        setattr(sys.modules[parent], child, m)


def load_reference():
    """Returns (TensorVMSplit, AlphaGridMask, LocalTensorfs, ray_utils module) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REF_ROOT}")
    _stub("kornia", create_meshgrid=lambda *a, **k: None)
    _stub("matplotlib", use=lambda *a, **k: None)
    _stub("matplotlib.pyplot")
    _stub("matplotlib.cm")
    _stub("plyfile")
    _stub("skimage")
    _stub("skimage.measure")
    _stub("imageio")
    _stub("lpips")
    if REF_PKG not in sys.path:
        sys.path.insert(0, REF_PKG)
    # the reference uses top-level names "models", "utils", "local_tensorfs"
    for clash in ("models", "utils", "local_tensorfs"):
        mod = sys.modules.get(clash)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REF_PKG):
            del sys.modules[clash]
    from models.tensoRF import TensorVMSplit  # noqa
    from models.tensorBase import AlphaGridMask  # noqa
    from local_tensorfs import LocalTensorfs  # noqa
    import utils.ray_utils as ray_utils  # noqa
    return TensorVMSplit, AlphaGridMask, LocalTensorfs, ray_utils
