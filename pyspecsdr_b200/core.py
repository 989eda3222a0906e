"""Host-side wrapper of the C ABI: a `Context` owns one pss_ctx on one GPU and exposes the batched
entry points for numpy arrays (host-pointer path, copies included) and for device pointers
(anything with `.data_ptr()` such as a torch CUDA tensor; asynchronous on the context's stream).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, filters
from ._lib import DemodDesc, DisplayOut, PipelineIO, PsdOut, PssError, lib

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


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class DemodPlan:
    """Device-resident filter tables for one (mode, sample_rate, block length)."""

    def __init__(self, ctx: "Context", mode: str, fs: float, N: int, iq_correct: bool = True):
        self.ctx, self.mode, self.fs, self.N = ctx, mode, float(fs), int(N)
        desc = DemodDesc()
        if mode not in filters.MODES:
            raise ValueError(f"no demodulation plan for mode {mode!r}")
        desc.mode = filters.MODES[mode]
        desc.N = N
        keep = []

        def arr(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            return _dp(a)

        if mode in ("NFM", "WFM"):
            p = filters.build_decim_plan(mode, fs, N)
            m = filters.build_modal_plan(p)
            self.host_plan, self.modal_plan = p, m
            desc.kind = 0
            for k in ("q", "n_out", "lead", "SF", "SB", "n_body", "m_tail", "tail_start", "tail_len"):
                setattr(desc, k, int(getattr(p, k)))
            desc.scale, desc.norm, desc.DB = np.float32(p.scale), p.norm, m.DB
            desc.iq_correct = 1 if (mode == "WFM" and iq_correct) else 0
            desc.body, desc.BF, desc.BB, desc.G = arr(m.body), arr(m.BF), arr(m.BB), arr(m.G)
            desc.CR, desc.CB = arr(m.CR), arr(m.CB)
            desc.head, desc.tail_T, desc.tail_M = arr(m.head), arr(m.tail_T), arr(m.tail_M)
        elif mode in ("USB", "LSB"):
            desc.kind = 1
            taps = filters.ssb_taps(fs)
            desc.taps, desc.n_taps = arr(taps), len(taps)
        elif mode == "AM":
            desc.kind = 2
            sos = filters.am_sos()
            desc.sos, desc.n_sections = arr(sos), len(sos)
        elif mode == "RAW":
            desc.kind = 3
        else:
            raise ValueError(f"no demodulation plan for mode {mode!r}")
        h = C.c_void_p()
        ctx._ck(lib.pss_demod_plan_create(ctx._h, C.byref(desc), C.byref(h)), f"pss_demod_plan_create({mode})")
        self._h = h
        self.out_len = int(lib.pss_demod_plan_out_len(h))
        self.channels = int(lib.pss_demod_plan_channels(h))

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            lib.pss_demod_plan_destroy(self.ctx._h, self._h)
        self._h = None


class Context:
    def __init__(self, device: int = 0):
        h = C.c_void_p()
        rc = lib.pss_init(device, C.byref(h))
        if rc != 0:
            raise PssError(f"pss_init(device={device}): {lib.pss_strerror(rc).decode()}")
        self._h = h
        self.device = device
        self._plans = {}
        self._streams = {}

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            for p in self._plans.values():
                p.close()
            self._plans.clear()
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
    def psd(self, samples, window="hamming", epilogue=False, W=0, want_stats=False, want_db=True, fp32=False):
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
        self._ck(lib.pss_psd_c64(self._h, x.ctypes.data, N, F, WINDOWS[window], 1 if epilogue else 0,
                                 1 if fp32 else 0, C.byref(out)), "pss_psd_c64")
        return res

    def psd_dev(self, iq, N, n_frames, db=None, window="hamming", epilogue=False, cols=None, W=0, stats=None,
                fp32=False, moments=None):
        """Device-pointer PSD: enqueue only.  `iq`, `db`, `cols`, `stats`, `moments` are device buffers."""
        out = PsdOut(_ptr(db), _ptr(cols), W, _ptr(stats), _ptr(moments))
        self._ck(lib.pss_psd_c64_dev(self._h, _ptr(iq), N, n_frames, WINDOWS[window], 1 if epilogue else 0,
                                     1 if fp32 else 0, C.byref(out)), "pss_psd_c64_dev")

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

    # ------------------------------------------------------------------ demodulation
    def demod_plan(self, mode: str, fs: float, N: int, iq_correct: bool = True) -> DemodPlan:
        """`iq_correct` (WFM only): True = demodulate_signal's chain (iq_correction fused into the kernel),
        False = demodulate_wfm called directly on whatever the caller passes."""
        key = (mode, float(fs), int(N), bool(iq_correct) or mode != "WFM")
        if key not in self._plans:
            self._plans[key] = DemodPlan(self, mode, fs, N, iq_correct)
        return self._plans[key]

    def demod(self, samples, fs: float, mode: str, iq_correct: bool = True) -> np.ndarray:
        """Batched demodulate_signal for one mode.  Returns float32 [F, out_len, channels]
        (channels = 2 for NFM/WFM, 1 for AM/USB/LSB/RAW)."""
        x = _as_frames(samples)
        F, N = x.shape
        plan = self.demod_plan(mode, fs, N, iq_correct)
        out = np.empty((F, plan.out_len, plan.channels), np.float32)
        self._ck(lib.pss_demod_c64(self._h, plan._h, x.ctypes.data, F, out.ctypes.data), f"pss_demod_c64({mode})")
        return out

    def demod_dev(self, plan: DemodPlan, iq, n_frames: int, audio, moments=None, frames_per_block=0):
        """Device-pointer demodulation.  `moments` = the PSD call's per-frame I/Q moments over the same
        IQ (frames_per_block rows per block, each over N / frames_per_block samples): lets a WFM plan skip
        its own iq_correction pass."""
        if moments is None:
            self._ck(lib.pss_demod_c64_dev(self._h, plan._h, _ptr(iq), n_frames, _ptr(audio)), "pss_demod_c64_dev")
        else:
            if frames_per_block < 1 or plan.N % frames_per_block:
                raise ValueError("frames_per_block must divide the plan's block length")
            self._ck(lib.pss_demod_c64_dev_moments(self._h, plan._h, _ptr(iq), n_frames, _ptr(audio), _ptr(moments),
                                                   frames_per_block, plan.N // frames_per_block),
                     "pss_demod_c64_dev_moments")

    # ------------------------------------------------------------------ display accumulate
    def display_render(self, cols, stats, rows_max=30, first=0, step=1, n_renders=None, guard_zero_range=False):
        """Waterfall / persistence history normalisation from the PSD kernel's `cols` and `stats`.
        Returns (norm [R, rows_max, W] newest row first, minmax [R, 2])."""
        cols = np.ascontiguousarray(cols, np.float32)
        stats = np.ascontiguousarray(stats, np.float32)
        F, W = cols.shape
        if n_renders is None:
            n_renders = (F - 1 - first) // max(step, 1) + 1
        norm = np.empty((n_renders, rows_max, W), np.float32)
        mm = np.empty((n_renders, 2), np.float32)
        self._ck(lib.pss_display_render(self._h, cols.ctypes.data, stats.ctypes.data, W, F, rows_max, first, step,
                                        n_renders, 1 if guard_zero_range else 0, norm.ctypes.data,
                                        mm.ctypes.data), "pss_display_render")
        return norm, mm

    def display_render_dev(self, cols, stats, W, n_frames, norm, minmax, rows_max=30, first=0, step=1,
                           n_renders=1, guard_zero_range=False):
        self._ck(lib.pss_display_render_dev(self._h, _ptr(cols), _ptr(stats), W, n_frames, rows_max, first, step,
                                            n_renders, 1 if guard_zero_range else 0, _ptr(norm), _ptr(minmax)),
                 "pss_display_render_dev")

    # ------------------------------------------------------------------ stateful display streams
    def display_open(self, stream: int, kind: str = "waterfall", W: int = 200, rows_max: int = 30, H: int = 0):
        """Create / reset display stream `stream`: the device-resident ring that plays the role of the
        reference's global WATERFALL_HISTORY / PERSISTENCE_HISTORY lists (pyspecsdr.py:130-131, 151-153)."""
        self._ck(lib.pss_display_open(self._h, stream, self.QUANT[kind], W, rows_max, H), "pss_display_open")
        self._streams[stream] = (kind, W, rows_max, H)

    def display_close(self, stream: int):
        self._ck(lib.pss_display_close(self._h, stream), "pss_display_close")
        self._streams.pop(stream, None)

    def display_rows(self, stream: int) -> int:
        return int(lib.pss_display_rows(self._h, stream))

    def display_accumulate(self, stream: int, rows, want=("norm64", "minmax64", "plane_a", "plane_b", "n_rows")):
        """Feed fp64 dB rows [n_rows, n_bins] (host) to a display stream, one render after every row - the
        call pattern of draw_waterfall / draw_gradient_waterfall / draw_persistence / draw_surface_plot.
        fp64 in numpy's operation order: the planes equal what the reference draws from the same rows.
        Returns dict(norm64 [n, rows_max, W], minmax64 [n, 2], plane_a, plane_b uint8 [n, rows_max, W],
        n_rows [n]); newest row first, NaN / 255 beyond the history."""
        kind, W, R, H = self._streams[stream]
        x = np.ascontiguousarray(rows, dtype=np.float64)
        if x.ndim == 1:
            x = x[None, :]
        n, nb = x.shape
        shapes = {"norm": ((n, R, W), np.float32), "norm64": ((n, R, W), np.float64), "minmax64": ((n, 2), np.float64),
                  "plane_a": ((n, R, W), np.uint8), "plane_b": ((n, R, W), np.uint8), "n_rows": ((n,), np.int32)}
        res = {k: np.empty(*shapes[k]) for k in want}
        out = DisplayOut(**{k: v.ctypes.data for k, v in res.items()})
        self._ck(lib.pss_display_accumulate_f64(self._h, stream, x.ctypes.data, nb, n, C.byref(out)),
                 "pss_display_accumulate_f64")
        if kind == "surface" and "n_rows" in res:
            res["n_rows"][:] = 1
        return res

    def display_accumulate_dev(self, stream: int, cols, stats, n_frames, first=0, step=1, n_renders=None, norm=None,
                               minmax=None, plane_a=None, plane_b=None, n_rows=None):
        """Device float32 rows from the PSD kernel (`cols`, `stats`) into a display stream; enqueue only."""
        if n_renders is None:
            n_renders = (n_frames - 1 - first) // max(step, 1) + 1
        out = DisplayOut(norm=_ptr(norm), minmax=_ptr(minmax), plane_a=_ptr(plane_a), plane_b=_ptr(plane_b),
                         n_rows=_ptr(n_rows))
        self._ck(lib.pss_display_accumulate_dev(self._h, stream, _ptr(cols), _ptr(stats), n_frames, first, step,
                                                n_renders, C.byref(out)), "pss_display_accumulate_dev")

    # ------------------------------------------------------------------ batched main-loop iteration
    def pipeline(self, blocks, fs: float, mode: str = "WFM", n_fft: int = 4096, W: int = 200, rows_max: int = 30,
                 want_db: bool = False, out=None, display_stream: int = -1, want_planes: bool = False):
        """Host buffers in, host buffers out: demodulate_signal + compute_fft/epilogue on every
        `n_fft` frame + waterfall accumulate after every block, for a batch of blocks.
        `blocks` is complex64 [n_blocks, N_block] (a pinned buffer makes the copies fast).
        Returns dict(audio, cols, stats, norm, minmax[, db][, plane_a, plane_b]).  `out` may supply
        preallocated arrays (checked: float32 / uint8, C-contiguous, exact shape).
        `display_stream` >= 0: a stream from display_open(); its history is carried across calls."""
        x = blocks if (isinstance(blocks, np.ndarray) and blocks.dtype == np.complex64 and blocks.ndim == 2
                       and blocks.flags.c_contiguous) else _as_frames(blocks)
        nb, N = x.shape
        fpb = N // n_fft
        plan = self.demod_plan(mode, fs, N) if mode else None
        o = out if out is not None else {}
        def buf(name, shape, dtype=np.float32):
            if name not in o:
                o[name] = np.empty(shape, dtype)
            a = o[name]
            if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.shape == tuple(shape) and a.flags.c_contiguous):
                raise ValueError(f"out[{name!r}] must be a C-contiguous {np.dtype(dtype).name} array of shape {tuple(shape)}")
            return a
        if display_stream >= 0:
            if display_stream not in self._streams:
                raise PssError(f"display stream {display_stream} is not open")
            _, sw, sr, _ = self._streams[display_stream]
            if (sw, sr) != (W, rows_max):
                raise ValueError("display stream geometry differs from W / rows_max")
        io = PipelineIO(N, n_fft, W, rows_max, plan._h if plan else None)
        io.display_stream = display_stream
        if want_planes:
            io.plane_a = buf("plane_a", (nb, rows_max, W), np.uint8).ctypes.data
            io.plane_b = buf("plane_b", (nb, rows_max, W), np.uint8).ctypes.data
        if plan:
            io.audio = buf("audio", (nb, plan.out_len, plan.channels)).ctypes.data
        io.cols = buf("cols", (nb * fpb, W)).ctypes.data
        io.stats = buf("stats", (nb * fpb, 4)).ctypes.data
        io.norm = buf("norm", (nb, rows_max, W)).ctypes.data
        io.minmax = buf("minmax", (nb, 2)).ctypes.data
        if want_db:
            io.db = buf("db", (nb * fpb, n_fft - 4)).ctypes.data
        self._ck(lib.pss_pipeline_c64(self._h, x.ctypes.data, nb, C.byref(io)), "pss_pipeline_c64")
        return o

    @staticmethod
    def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
        """numpy array over page-locked memory from pss_host_alloc (freed when the array dies)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = lib.pss_host_alloc(max(n, 1))
        if not p:
            raise PssError("pss_host_alloc failed")
        raw = (C.c_char * max(n, 1)).from_address(p)
        arr = np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        import weakref
        weakref.finalize(raw, lib.pss_host_free, p)
        return arr

    # ------------------------------------------------------------------ helpers either side of the path
    def iq_correct(self, samples) -> np.ndarray:
        """iq_correction (signal_processing.py:46-80) -> complex64, same shape as the input."""
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        fr = x[None, :] if x.ndim == 1 else x
        out = np.empty(fr.shape, np.complex64)
        self._ck(lib.pss_iq_correct_c64(self._h, fr.ctypes.data, fr.shape[1], fr.shape[0], out.ctypes.data),
                 "pss_iq_correct_c64")
        return out[0] if x.ndim == 1 else out

    def sosfilt(self, sos, data) -> np.ndarray:
        """scipy.signal.sosfilt(sos, data) with zero initial state; float64 result like scipy."""
        d = np.ascontiguousarray(data, dtype=np.float32)
        fr = d[None, :] if d.ndim == 1 else d
        sos = np.ascontiguousarray(sos, dtype=np.float64)
        y = np.empty(fr.shape, np.float32)
        self._ck(lib.pss_sosfilt_f32(self._h, fr.ctypes.data, fr.shape[1], fr.shape[0], sos.ctypes.data, len(sos),
                                     y.ctypes.data), "pss_sosfilt_f32")
        y = y.astype(np.float64)
        return y[0] if d.ndim == 1 else y

    def bandpass(self, data, lowcut, highcut, sample_rate) -> np.ndarray:
        """bandpass_filter (signal_processing.py:34-42): design on the host with scipy, filter on the GPU."""
        return self.sosfilt(filters.butter_sos(lowcut, highcut, sample_rate), data)

    def signal_power(self, samples) -> np.ndarray:
        """measure_signal_power (signal_processing.py:325-328) per block -> float32 dB."""
        x = _as_frames(samples)
        out = np.empty(len(x), np.float32)
        self._ck(lib.pss_power_c64(self._h, x.ctypes.data, x.shape[1], x.shape[0], out.ctypes.data), "pss_power_c64")
        return out

    CLASS_LABELS = ("UNKNOWN", "FM_BROADCAST", "NARROW_FM", "AM_BROADCAST", "SSB", "DIGITAL")

    def classify(self, samples, fs: float):
        """classify_signal's computation (signal_processing.py:296-322 with the missing `welch` import
        supplied) per block -> (labels [list of str], features float64 [n_blocks, 4] = signal_bw,
        modulation_index, spectral_flatness, Welch peak dB)."""
        x = _as_frames(samples)
        feat = np.empty((len(x), 4), np.float64)
        lab = np.empty(len(x), np.int32)
        self._ck(lib.pss_classify_c64(self._h, x.ctypes.data, x.shape[1], x.shape[0], float(fs), feat.ctypes.data,
                                      lab.ctypes.data), "pss_classify_c64")
        return [self.CLASS_LABELS[i] for i in lab], feat

    def classify_dev(self, iq, N, n_blocks, fs, features, label):
        self._ck(lib.pss_classify_c64_dev(self._h, _ptr(iq), N, n_blocks, float(fs), _ptr(features), _ptr(label)),
                 "pss_classify_c64_dev")

    def to_int16(self, audio) -> np.ndarray:
        """write_audio_samples' numeric line (audio_processing.py:36-38): np.int16(samples * 32767).
        float64 input (what the reference holds there) takes the fp64 entry point: bit-identical to numpy."""
        a = np.asarray(audio)
        out = np.empty(a.shape, np.int16)
        if a.dtype == np.float32:
            a = np.ascontiguousarray(a)
            self._ck(lib.pss_audio_to_int16(self._h, a.ctypes.data, a.size, out.ctypes.data), "pss_audio_to_int16")
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
            self._ck(lib.pss_audio_to_int16_f64(self._h, a.ctypes.data, a.size, out.ctypes.data),
                     "pss_audio_to_int16_f64")
        return out

    def spectrum_normalise(self, db_rows, W: int):
        """draw_spectrogram's numeric part (pyspecsdr.py:418-452).  Returns (cols [F, W] in [0,1],
        range [F, 2] = display_min, display_max)."""
        d = np.asarray(db_rows)
        dt, fn = (np.float64, lib.pss_spectrum_normalise_f64) if d.dtype == np.float64 else (np.float32, lib.pss_spectrum_normalise)
        d = np.ascontiguousarray(d, dtype=dt)
        fr = d[None, :] if d.ndim == 1 else d
        cols = np.empty((len(fr), W), dt)
        rng = np.empty((len(fr), 2), dt)
        self._ck(fn(self._h, fr.ctypes.data, fr.shape[1], fr.shape[0], W, cols.ctypes.data, rng.ctypes.data),
                 "pss_spectrum_normalise")
        return cols, rng

    def surface_row(self, cols, stats):
        """draw_surface_plot's numeric part (pyspecsdr.py:1575-1596) from a PSD row's W-column resample
        and its min/max: magnitude = int(normalised * 20)."""
        norm, mm = self.display_render(cols, stats, rows_max=1, guard_zero_range=True)
        return (norm[:, 0].astype(np.float64) * 20).astype(np.int64), mm

    QUANT = {"waterfall": 0, "gradient": 1, "persistence": 2, "surface": 3}

    def display_quantise(self, norm, kind: str = "waterfall", H: int = 0):
        """Glyph / colour index planes (uint8, 255 = no data) of the draw_* functions from normalised
        values; see pss_display_quantise in include/pss.h."""
        v = np.ascontiguousarray(norm, dtype=np.float32)
        a = np.empty(v.shape, np.uint8)
        b = np.empty(v.shape, np.uint8)
        self._ck(lib.pss_display_quantise(self._h, v.ctypes.data, v.size, self.QUANT[kind], H, a.ctypes.data,
                                          b.ctypes.data), "pss_display_quantise")
        return a, b
