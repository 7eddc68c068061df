"""Drop-in for the reference's `local_tensorfs` module name (train.py:21, renderer.py)."""
from localrf_b200.local_tensorfs import LocalTensorfs, ids2pixel, ids2pixel_view  # noqa: F401
