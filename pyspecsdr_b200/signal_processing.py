"""Drop-in replacement for PySpecSDR's `signal_processing` module.

Put this package directory ahead of the reference on `sys.path` and
`from signal_processing import *` in an unmodified pyspecsdr.py (line 98) binds these functions.
Same names, argument meaning, return shapes and dtypes as /root/reference/signal_processing.py;
the numerics run in libpss.so on the GPU (there is no CPU fallback: a missing library or GPU raises).

Return-dtype note: the kernels produce float32 on the wire (the consumer is a float32 PortAudio
stream, pyspecsdr.py:2341-2348); results are widened to float64 here so callers see the dtype the
reference returns.
"""
from __future__ import annotations

import numpy as np
# names the reference module leaks through `import *` (pyspecsdr.py relies on `np` at least)
from scipy.signal import butter, lfilter            # noqa: F401
from scipy.signal import firwin                     # noqa: F401
from scipy.signal import hilbert                    # noqa: F401
from scipy.signal import decimate                   # noqa: F401
from scipy.signal import bilinear                   # noqa: F401
from scipy.signal import resample_poly              # noqa: F401

from . import core as _core
from .filters import AUDIO_RATE as DEFAULT_SAMPLE_RATE, BUTTER_ORDER   # pyspecconst.py:3,5

_CTX = None


def _ctx() -> "_core.Context":
    global _CTX
    if _CTX is None:
        import os
        _CTX = _core.Context(int(os.environ.get("PSS_DEVICE", "0")))
    return _CTX


def _c64(samples) -> np.ndarray:
    return np.ascontiguousarray(samples, dtype=np.complex64)


def compute_fft(samples):
    """signal_processing.py:243-264 -> float64 [N] dB, fft-shifted, Hamming window."""
    x = _c64(samples)
    return _ctx().psd(x, window="hamming")["db"][0].astype(np.float64)


def mono_to_stereo(mono_audio):
    """signal_processing.py:83-88."""
    m = np.asarray(mono_audio, dtype=np.float64)
    return np.repeat(m[:, None], 2, axis=1)


def iq_correction(samples: np.ndarray) -> np.ndarray:
    """signal_processing.py:46-80.  The I channel comes from the GPU RAW plan; the full complex
    result is only needed by callers outside the hot path, so Q is rebuilt with the same float32
    arithmetic from the same block moments (one extra pass on the host is avoided by asking the
    library for both channels)."""
    x = _c64(samples)
    return _ctx().iq_correct(x)


def demodulate_nfm(samples, sample_rate, target_rate=DEFAULT_SAMPLE_RATE):
    """signal_processing.py:91-116 -> float64 [ceil((N-1)/q), 2]."""
    _check_rate(target_rate)
    return _ctx().demod(_c64(samples), float(sample_rate), "NFM")[0].astype(np.float64)


def demodulate_wfm(samples, sample_rate, target_rate=DEFAULT_SAMPLE_RATE):
    """signal_processing.py:119-176: the WFM chain on the samples as given, NO iq_correction - exactly
    like the reference's function (its dispatcher corrects first, :222-225; that fused path is
    demodulate_signal(..., 'WFM'))."""
    _check_rate(target_rate)
    return _ctx().demod(_c64(samples), float(sample_rate), "WFM", iq_correct=False)[0].astype(np.float64)


def demodulate_am(samples):
    """signal_processing.py:179-195 -> float64 [N, 2]."""
    return mono_to_stereo(_ctx().demod(_c64(samples), float(DEFAULT_SAMPLE_RATE), "AM")[0, :, 0])


def demodulate_ssb(samples, sample_rate, lower=True):
    """signal_processing.py:198-217 -> float64 [N, 2]; `lower` has no effect in the reference either."""
    return mono_to_stereo(_ctx().demod(_c64(samples), float(sample_rate), "LSB" if lower else "USB")[0, :, 0])


def demodulate_signal(samples, sample_rate, mode='NFM'):
    """signal_processing.py:220-240: dispatch, iq_correction for the non-voice modes."""
    x = _c64(samples)
    fs = float(sample_rate)
    if mode == 'NFM':
        return _ctx().demod(x, fs, "NFM")[0].astype(np.float64)
    if mode == 'WFM':
        return _ctx().demod(x, fs, "WFM")[0].astype(np.float64)      # correction fused in the kernel
    if mode == 'AM':
        return demodulate_am(x)
    if mode == 'USB':
        return demodulate_ssb(x, fs, lower=False)
    if mode == 'LSB':
        return demodulate_ssb(x, fs, lower=True)
    if mode == 'RAW':
        return _ctx().demod(x, fs, "RAW")[0, :, 0]                   # float32 1-D, like the reference
    return np.zeros((len(x), 2))                                     # :240


def measure_signal_power(samples):
    """signal_processing.py:325-328 -> float (dB)."""
    return float(_ctx().signal_power(_c64(samples))[0])


def bandpass_filter(data, lowcut, highcut, sample_rate):
    """signal_processing.py:34-42 (Butterworth order BUTTER_ORDER, forward sosfilt, zero state)."""
    return _ctx().bandpass(np.asarray(data), float(lowcut), float(highcut), float(sample_rate))


ENABLE_CLASSIFIER = False      # opt-in (or PSS_CLASSIFIER=1): see classify_signal


def classify_signal(samples, sample_rate, bandwidth):
    """signal_processing.py:296-322.  In the reference this raises NameError (`welch` is never
    imported, :299) and the callers catch it (pyspecsdr.py:2571), so by default it keeps failing the
    same way.  With `ENABLE_CLASSIFIER = True` (or PSS_CLASSIFIER=1 in the environment) the intended
    computation — scipy-default Welch PSD, estimate_bandwidth, estimate_modulation_index, spectral
    flatness, the decision tree — runs on the GPU (`pss_classify_c64`); checked against the reference
    executed with that one name supplied (tests/golden/classifier.npz)."""
    import os
    if not (ENABLE_CLASSIFIER or os.environ.get("PSS_CLASSIFIER") == "1"):
        raise NameError("name 'welch' is not defined")
    labels, _ = _ctx().classify(_c64(samples), float(sample_rate))
    return labels[0]


def _check_rate(target_rate):
    if target_rate != DEFAULT_SAMPLE_RATE:
        raise ValueError("only the reference's DEFAULT_SAMPLE_RATE (22050 Hz) audio rate is supported")
