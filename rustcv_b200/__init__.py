"""rustcv_b200 -- B200 (sm_100a) imgproc backend for RustCV's per-pixel hot path.

Host-side mirror of the reference API (`Mat`, `imgproc::*`, `videoio` decode dispatch)
over the C ABI of include/rcv_imgproc.h.  Importing the package loads
rustcv_b200/librcv_imgproc.so; there is no CPU fallback.
"""
from . import _ffi  # noqa: F401  (raises ImportError when the library is not built)
from . import imgproc, videoio  # noqa: F401
from .mat import F32, U8, Mat, MatBatch  # noqa: F401

__all__ = ["Mat", "MatBatch", "U8", "F32", "imgproc", "videoio"]
