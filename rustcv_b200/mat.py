"""`Mat` -- host-side mirror of rustcv::core::Mat (rustcv/src/core/mat.rs:6-51) with the
storage variants BASELINE.json's north_star adds: device-resident (HBM) and pinned host.

Field names and meaning follow the reference: `rows`, `cols`, `step` (bytes per row,
>= cols*channels*elemsize), `channels`; `data` is a flat u8 buffer for host Mats.  The
reference Mat is u8-only (TODO at mat.rs:53); `depth` tags f32 images, whose `step` stays
in bytes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F

U8, F32 = F.RCV_U8, F.RCV_F32
_NP = {U8: np.uint8, F32: np.float32}


def _elem(depth: int) -> int:
    return 4 if depth == F32 else 1


class Mat:
    __slots__ = ("data", "rows", "cols", "step", "channels", "depth", "loc", "device", "_pinned_ptr", "_owner",
                 "_registered")

    def __init__(self):
        self.data = None  # np.uint8 1-D for host Mats, int (device pointer) for device Mats
        self.rows = 0
        self.cols = 0
        self.step = 0
        self.channels = 0
        self.depth = U8
        self.loc = F.RCV_HOST
        self.device = 0
        self._pinned_ptr = None
        self._owner = None  # the batch allocation a device Mat was carved from
        self._registered = None  # address page-locked in place by register()

    # -- constructors (mat.rs:18-40) ------------------------------------------------
    @staticmethod
    def new(rows: int, cols: int, channels: int, depth: int = U8, step: int | None = None) -> "Mat":
        """Mat::new: packed (`step = cols*channels`) zero-filled host Mat; `step` may pad."""
        m = Mat()
        rb = cols * channels * _elem(depth)
        m.rows, m.cols, m.channels, m.depth = rows, cols, channels, depth
        m.step = rb if step is None else step
        assert m.step >= rb
        m.data = np.zeros(rows * m.step, dtype=np.uint8)
        return m

    @staticmethod
    def empty() -> "Mat":
        return Mat()

    @staticmethod
    def from_numpy(a: np.ndarray) -> "Mat":
        """Wraps (copies) rows x cols [x channels] u8/f32 into a packed host Mat."""
        a = np.ascontiguousarray(a)
        depth = F32 if a.dtype == np.float32 else U8
        assert a.dtype in (np.uint8, np.float32) and a.ndim in (2, 3)
        cn = 1 if a.ndim == 2 else a.shape[2]
        m = Mat.new(a.shape[0], a.shape[1], cn, depth)
        m.data[:] = a.view(np.uint8).ravel()
        return m

    @staticmethod
    def from_numpy_strided(a: np.ndarray, step: int, fill: int = 0xA5) -> "Mat":
        """Same, with `step` > row bytes (padding bytes set to `fill`)."""
        a = np.ascontiguousarray(a)
        depth = F32 if a.dtype == np.float32 else U8
        cn = 1 if a.ndim == 2 else a.shape[2]
        m = Mat.new(a.shape[0], a.shape[1], cn, depth, step=step)
        m.data[:] = fill
        rb = m.cols * m.channels * _elem(depth)
        m.data.reshape(m.rows, step)[:, :rb] = a.view(np.uint8).reshape(m.rows, rb)
        return m

    @staticmethod
    def pinned(rows: int, cols: int, channels: int, depth: int = U8, device: int = -1) -> "Mat":
        """Host Mat in page-locked memory (rcv_pinned_alloc_on: on the NUMA node of GPU `device`
        when the box has several) for direct DMA."""
        m = Mat()
        m.rows, m.cols, m.channels, m.depth = rows, cols, channels, depth
        m.step = cols * channels * _elem(depth)
        n = max(rows * m.step, 1)
        p = C.c_void_p()
        F.check(F.lib.rcv_pinned_alloc_on(device, C.byref(p), n))
        m._pinned_ptr = p.value
        m.data = np.ctypeslib.as_array((C.c_uint8 * n).from_address(p.value))[: rows * m.step]
        m.loc = F.RCV_HOST_PINNED
        return m

    @staticmethod
    def device_new(rows: int, cols: int, channels: int, depth: int = U8, device: int = -1) -> "Mat":
        """Device-resident Mat (HBM), step rounded up to 256 B."""
        c = F.RcvMat()
        F.check(F.lib.rcv_mat_alloc_device(C.byref(c), rows, cols, channels, depth, device))
        return Mat._from_c(c, owner=None)

    @staticmethod
    def device_batch(n: int, rows: int, cols: int, channels: int, depth: int = U8, device: int = -1) -> "MatBatch":
        arr = (F.RcvMat * n)()
        F.check(F.lib.rcv_mat_alloc_device_batch(arr, n, rows, cols, channels, depth, device))
        return MatBatch(arr, n, owned=True)

    @staticmethod
    def _from_c(c: F.RcvMat, owner) -> "Mat":
        m = Mat()
        m.data = c.data
        m.rows, m.cols, m.step = c.rows, c.cols, c.step
        m.channels, m.depth, m.loc, m.device = c.channels, c.depth, c.loc, c.device
        m._owner = owner
        return m

    # -- reference API -----------------------------------------------------------------
    def is_empty(self) -> bool:
        """mat.rs:42-44"""
        if self.loc == F.RCV_DEVICE:
            return not self.data or self.rows == 0 or self.cols == 0
        return self.data is None or self.data.size == 0 or self.rows == 0 or self.cols == 0

    def row_bytes(self, row: int) -> np.ndarray:
        """mat.rs:47-51: the valid bytes of a row, padding dropped."""
        assert self.loc != F.RCV_DEVICE
        start = row * self.step
        return self.data[start:start + self.cols * self.channels * _elem(self.depth)]

    def ensure_size(self, rows: int, cols: int, channels: int, depth: int = U8) -> None:
        """rustcv-camera/src/mat.rs:65-74 / videoio/mod.rs:192-199: (re)size a host dst the
        way `read` does -- packed, reallocating only when the byte length changes.  A buffer
        that was page-locked in place is unregistered before it is dropped and the new one
        registered again (what the Rust wrapper does)."""
        assert self.loc == F.RCV_HOST
        step = cols * channels * _elem(depth)
        if self.data is None or self.data.size != rows * step:
            was_registered = self._registered is not None
            self.unregister()
            self.data = np.zeros(rows * step, dtype=np.uint8)
            if was_registered:
                self.register()
        self.rows, self.cols, self.channels, self.depth, self.step = rows, cols, channels, depth, step

    def register(self) -> "Mat":
        """Page-locks this host Mat's buffer in place (rcv_host_register): the reference reuses
        the Vec<u8> frame after frame, so one registration serves every later call."""
        assert self.loc == F.RCV_HOST
        if self._registered is None and self.data is not None and self.data.size:
            F.check(F.lib.rcv_host_register(self.data.ctypes.data, self.data.size))
            self._registered = self.data.ctypes.data
        return self

    def unregister(self) -> None:
        if self._registered is not None:
            p, self._registered = self._registered, None
            F.check(F.lib.rcv_host_unregister(p))

    # -- conversions ----------------------------------------------------------------------
    def to_numpy(self) -> np.ndarray:
        """rows x cols [x channels] array (copy); device Mats are downloaded."""
        if self.loc == F.RCV_DEVICE:
            h = Mat.new(self.rows, self.cols, self.channels, self.depth)
            if self.rows and self.cols:
                F.check(F.lib.rcv_mat_download(C.byref(self.c()), C.byref(h.c())))
            return h.to_numpy()
        rb = self.cols * self.channels * _elem(self.depth)
        rows = self.data.reshape(self.rows, self.step)[:, :rb] if self.rows else self.data.reshape(0, 0)
        a = np.ascontiguousarray(rows).view(_NP[self.depth])
        shape = (self.rows, self.cols) if self.channels == 1 else (self.rows, self.cols, self.channels)
        return a.reshape(shape).copy()

    def upload(self, device: int = -1) -> "Mat":
        d = Mat.device_new(self.rows, self.cols, self.channels, self.depth, device)
        if self.rows and self.cols:
            F.check(F.lib.rcv_mat_upload(C.byref(self.c()), C.byref(d.c())))
        return d

    def like(self, channels: int | None = None, depth: int | None = None, rows: int | None = None,
             cols: int | None = None) -> "Mat":
        """A fresh Mat of this one's location with (optionally) different geometry."""
        r = self.rows if rows is None else rows
        c = self.cols if cols is None else cols
        cn = self.channels if channels is None else channels
        dp = self.depth if depth is None else depth
        if self.loc == F.RCV_DEVICE:
            return Mat.device_new(r, c, cn, dp, self.device)
        if self.loc == F.RCV_HOST_PINNED:
            return Mat.pinned(r, c, cn, dp)
        return Mat.new(r, c, cn, dp)

    def c(self) -> F.RcvMat:
        """The POD handed across the C ABI (rebuilt each call: fields are public/mutable)."""
        c = F.RcvMat()
        if self.loc == F.RCV_DEVICE:
            c.data = self.data
        else:
            c.data = self.data.ctypes.data if self.data is not None and self.data.size else None
        c.rows, c.cols, c.step = self.rows, self.cols, self.step
        c.channels, c.depth, c.loc, c.device = self.channels, self.depth, self.loc, self.device
        # the POD borrows self.data: it keeps the owner alive, never the other way round (a Mat holding its own
        # POD would be a reference cycle, and device / pinned storage would wait for the cyclic GC)
        c._keep = self
        return c

    def free(self) -> None:
        if self._registered is not None:
            self.unregister()
        if self.loc == F.RCV_DEVICE and self.data and self._owner is None:
            c = self.c()
            F.check(F.lib.rcv_mat_free_device(C.byref(c)))
            self.data = None
        elif self._pinned_ptr:
            self.data = None
            F.check(F.lib.rcv_pinned_free(self._pinned_ptr))
            self._pinned_ptr = None

    def __del__(self):  # device / pinned storage is owned by the Mat (a Rust Drop would call the same frees)
        try:
            self.free()
        except Exception:  # noqa: BLE001  (interpreter shutdown, context already gone)
            pass

    def __repr__(self) -> str:  # mat.rs:56-64
        return (f"Mat {{ rows: {self.rows}, cols: {self.cols}, channels: {self.channels}, step: {self.step}, "
                f"depth: {self.depth}, loc: {self.loc} }}")


class MatBatch:
    """n device Mats of one geometry in ONE allocation (rcv_mat_alloc_device_batch), or a
    plain list of Mats presented as a C array."""

    def __init__(self, arr, n: int, owned: bool, mats: list | None = None):
        self.arr, self.n, self.owned = arr, n, owned
        self.mats = mats if mats is not None else [Mat._from_c(arr[i], owner=self) for i in range(n)]

    @staticmethod
    def of(mats: list) -> "MatBatch":
        arr = (F.RcvMat * len(mats))()
        for i, m in enumerate(mats):
            arr[i] = m.c()
        return MatBatch(arr, len(mats), owned=False, mats=list(mats))

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int) -> Mat:
        return self.mats[i]

    def free(self) -> None:
        if self.owned and self.n:
            self.owned = False
            for m in self.mats:
                m.data = None
            F.check(F.lib.rcv_mat_free_device_batch(self.arr, self.n))

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001
            pass
