"""Drop-in for `models.tensorBase` (local_tensorfs.py:6)."""
from localrf_b200.tensorf import AlphaGridMask, MLPRender_Fea_late_view, TensorBase  # noqa: F401
