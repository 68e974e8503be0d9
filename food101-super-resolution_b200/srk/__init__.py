"""srk - Python host side of libsrk, the B200 (sm_100a) kernels behind the SR hot path of
Jaskieeeer/food101-super-resolution (src/models.py, src/loss.py, src/metrics.py)."""
from . import _lib  # noqa: F401  (raises if libsrk.so is missing: there is no fallback)
from . import fn, ops  # noqa: F401
from .ops import cfg, set_compute_dtype, set_conv_impl, set_overlap_wgrad  # noqa: F401

__all__ = ["fn", "ops", "cfg", "set_compute_dtype", "set_conv_impl", "set_overlap_wgrad"]
from . import optim  # noqa: F401,E402
from . import data, dp, evaluate, trainer  # noqa: F401,E402
