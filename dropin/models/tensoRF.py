"""Drop-in for `models.tensoRF` (local_tensorfs.py:8)."""
from localrf_b200.tensorf import TensorVMSplit  # noqa: F401
