"""Host-side wrapper of the C ABI: a `Context` owns one pss_ctx on one GPU and exposes the batched
entry points for numpy arrays (host-pointer path, copies included) and for device pointers
(anything with `.data_ptr()` such as a torch CUDA tensor; asynchronous on the context's stream).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PsdOut, PssError, lib

WINDOWS = {"none": 0, "hamming": 1, "hann": 2, None: 0}


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()          # torch tensor (device or pinned host)


def _as_frames(samples) -> np.ndarray:
    x = np.ascontiguousarray(samples, dtype=np.complex64)
    if x.ndim == 1:
        x = x[None, :]
    if x.ndim != 2:
        raise ValueError("IQ must be [N] or [n_frames, N] complex64")
    return x


class Context:
    def __init__(self, device: int = 0):
        h = C.c_void_p()
        rc = lib.pss_init(device, C.byref(h))
        if rc != 0:
            raise PssError(f"pss_init(device={device}): {lib.pss_strerror(rc).decode()}")
        self._h = h
        self.device = device

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            lib.pss_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        _lib.check(self._h, rc, what)

    def set_stream(self, cuda_stream_handle):
        """Adopt a cudaStream_t (int), e.g. torch.cuda.current_stream().cuda_stream (0 = the CUDA
        default stream); None goes back to the context's own stream."""
        if cuda_stream_handle is None:
            self._ck(lib.pss_use_own_stream(self._h), "pss_use_own_stream")
        else:
            self._ck(lib.pss_set_stream(self._h, cuda_stream_handle or None), "pss_set_stream")

    def sync(self):
        self._ck(lib.pss_sync(self._h), "pss_sync")

    @property
    def launches(self) -> int:
        return int(lib.pss_kernel_launches(self._h))

    # ------------------------------------------------------------------ PSD (host arrays)
    def psd(self, samples, window="hamming", epilogue=False, W=0, want_stats=False, want_db=True):
        """Batched compute_fft (+ optional main-loop epilogue).  Returns dict of float32 arrays:
        db [F, N or N-4], cols [F, W], stats [F, 4] (max, mean, finite-min, finite-max)."""
        x = _as_frames(samples)
        F, N = x.shape
        n_out = N - 4 if epilogue else N
        res = {}
        out = PsdOut()
        if want_db:
            res["db"] = np.empty((F, n_out), np.float32)
            out.db = res["db"].ctypes.data
        if W:
            res["cols"] = np.empty((F, W), np.float32)
            out.cols = res["cols"].ctypes.data
            out.W = W
        if want_stats:
            res["stats"] = np.empty((F, 4), np.float32)
            out.stats = res["stats"].ctypes.data
        self._ck(lib.pss_psd_c64(self._h, x.ctypes.data, N, F, WINDOWS[window], 1 if epilogue else 0, 0,
                                 C.byref(out)), "pss_psd_c64")
        return res

    def psd_dev(self, iq, N, n_frames, db=None, window="hamming", epilogue=False, cols=None, W=0, stats=None):
        """Device-pointer PSD: enqueue only.  `iq`, `db`, `cols`, `stats` are device buffers."""
        out = PsdOut(_ptr(db), _ptr(cols), W, _ptr(stats))
        self._ck(lib.pss_psd_c64_dev(self._h, _ptr(iq), N, n_frames, WINDOWS[window], 1 if epilogue else 0, 0,
                                     C.byref(out)), "pss_psd_c64_dev")

    # ------------------------------------------------------------------ scanner
    def scan(self, frames, rel_db=20.0, threshold=None, want_rows=False):
        """Per-step un-windowed PSD peak and above-threshold bin count (pyspecsdr.py:2542-2552)."""
        x = _as_frames(frames)
        F, N = x.shape
        peak = np.empty(F, np.float32)
        count = np.empty(F, np.int32)
        rows = np.empty((F, N), np.float32) if want_rows else None
        use_abs, thr = (0, rel_db) if threshold is None else (1, threshold)
        self._ck(lib.pss_scan_c64(self._h, x.ctypes.data, N, F, use_abs, thr, peak.ctypes.data,
                                  count.ctypes.data, _ptr(rows)), "pss_scan_c64")
        return (peak, count, rows) if want_rows else (peak, count)

    def scan_dev(self, iq, N, n_steps, peak, count, rows=None, rel_db=20.0, threshold=None):
        use_abs, thr = (0, rel_db) if threshold is None else (1, threshold)
        self._ck(lib.pss_scan_c64_dev(self._h, _ptr(iq), N, n_steps, use_abs, thr, _ptr(peak), _ptr(count),
                                      _ptr(rows)), "pss_scan_c64_dev")
