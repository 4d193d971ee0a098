"""rustcv_b200 -- B200 (sm_100a) imgproc backend for RustCV's per-pixel hot path.

Host-side mirror of the reference API (`Mat`, `imgproc::*`, `videoio` decode dispatch)
over the C ABI of include/rcv_imgproc.h.  The first use of `Mat`, `imgproc`, `videoio`
(or `_ffi`) loads rustcv_b200/librcv_imgproc.so and raises ImportError if it is not
built -- there is no CPU fallback.  `python -m rustcv_b200.build` builds it (the package
itself imports without the library so that the build module is reachable on a clean tree).
"""
import importlib

__all__ = ["Mat", "MatBatch", "U8", "F32", "imgproc", "videoio"]

_LAZY = {"Mat": "mat", "MatBatch": "mat", "U8": "mat", "F32": "mat", "imgproc": "imgproc", "videoio": "videoio",
         "_ffi": "_ffi", "mat": "mat"}


def __getattr__(name):
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError(f"module 'rustcv_b200' has no attribute {name!r}")
    m = importlib.import_module(f".{mod}", __name__)  # raises ImportError when the .so is missing
    return m if name in ("imgproc", "videoio", "_ffi", "mat") else getattr(m, name)
