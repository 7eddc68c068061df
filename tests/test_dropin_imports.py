"""CPU-only: the dropin/ directory resolves the module names the reference's train.py / renderer.py
import (`local_tensorfs`, `models.tensoRF`, `models.tensorBase`) to this package."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_module_names_resolve_to_this_package():
    code = (
        "import sys; sys.path[:0] = [%r, %r];"
        "import local_tensorfs, models.tensoRF, models.tensorBase;"
        "import localrf_b200 as L;"
        "assert local_tensorfs.LocalTensorfs is L.LocalTensorfs;"
        "assert models.tensoRF.TensorVMSplit is L.TensorVMSplit;"
        "assert models.tensorBase.AlphaGridMask is L.AlphaGridMask;"
        "print('ok')" % (os.path.join(ROOT, "dropin"), ROOT))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
