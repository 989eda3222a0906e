"""Seeded synthetic IQ generators (numpy, complex64) for tests, golden vectors and bench.

These replace the SDR device read (`SDRDevice.read_samples`, reference
pyspecsdr.py:1885-1891 returns 1-D complex64) with deterministic signals. The
signal classes follow SURVEY.md §8(d): white noise, tone+noise (the case that
breaks an fp32 FFT), wide-band FM tone, AM tone, SSB two-tone and the scanner's
per-step carrier. Everything is float64 internally and rounded once to complex64,
exactly what a CF32 stream from SoapySDR would deliver.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "noise", "tone_noise", "wbfm", "am_tone", "ssb_two_tone", "scanner_frames", "halfband",
    "impulse", "KINDS", "make",
]


def _cnoise(rng: np.random.Generator, n: int, sigma: float) -> np.ndarray:
    # complex Gaussian with total power sigma**2 (sigma/sqrt(2) per component)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (sigma / np.sqrt(2.0))


def noise(n: int, seed: int = 0) -> np.ndarray:
    """White complex Gaussian, unit power."""
    rng = np.random.default_rng(seed)
    return _cnoise(rng, n, 1.0).astype(np.complex64)


def tone_noise(n: int, seed: int = 0, f: float = 0.1234, dbc: float = -40.0) -> np.ndarray:
    """Unit complex tone at `f` cycles/sample plus white noise `dbc` dB below it."""
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    x = np.exp(2j * np.pi * f * k) + _cnoise(rng, n, 10.0 ** (dbc / 20.0))
    return x.astype(np.complex64)


def wbfm(n: int, seed: int = 0, fs: float = 2.4e6, dev: float = 75e3, fm: float = 1e3,
         dbc: float = -40.0) -> np.ndarray:
    """Broadcast-style FM: a `fm` Hz tone at +-`dev` Hz deviation, plus noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    m = np.sin(2 * np.pi * fm * t + 0.3 * seed)
    phase = 2 * np.pi * dev * np.cumsum(m) / fs
    x = np.exp(1j * phase) + _cnoise(rng, n, 10.0 ** (dbc / 20.0))
    return x.astype(np.complex64)


def am_tone(n: int, seed: int = 0, fs: float = 1e6, fm: float = 1e3, depth: float = 0.5,
            dbc: float = -40.0) -> np.ndarray:
    """AM carrier at DC with a `fm` Hz tone at modulation depth `depth`, plus noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    phi0 = 0.7 + 0.1 * seed
    x = (1.0 + depth * np.sin(2 * np.pi * fm * t)) * np.exp(1j * phi0)
    x = x + _cnoise(rng, n, 10.0 ** (dbc / 20.0))
    return x.astype(np.complex64)


def ssb_two_tone(n: int, seed: int = 0, fs: float = 1e6, f1: float = 700.0, f2: float = 1900.0,
                 dbc: float = -40.0) -> np.ndarray:
    """Two-tone SSB test signal (both tones on the upper side), plus noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    x = np.exp(2j * np.pi * f1 * t) + np.exp(2j * np.pi * f2 * t)
    x = x + _cnoise(rng, n, 10.0 ** (dbc / 20.0))
    return x.astype(np.complex64)


def scanner_frames(n_steps: int, n: int, seed: int = 0) -> np.ndarray:
    """[n_steps, n] frames: noise plus a carrier whose level/offset/width depend on the step."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_steps, n), dtype=np.complex64)
    k = np.arange(n)
    for s in range(n_steps):
        level = 10.0 ** ((-30.0 + 35.0 * ((s * 7) % 11) / 10.0) / 20.0)
        off = ((s * 37) % 97) / 97.0 - 0.5          # cycles/sample in [-0.5, 0.5)
        width = 0.002 * (1 + (s % 5))               # FM-ish spreading
        ph = 2 * np.pi * (off * k + width * np.cumsum(np.sin(2 * np.pi * 0.001 * (1 + s % 3) * k)))
        x = level * np.exp(1j * ph) + _cnoise(rng, n, 0.05)
        out[s] = x.astype(np.complex64)
    return out


def halfband(n: int, seed: int = 0, frac: float = 0.6, floor_db: float = -35.0) -> np.ndarray:
    """Noise filling `frac` of the band at 0 dB and the rest `floor_db` lower: more than half of the
    spectrum sits far above the remaining bins, so the main loop's median-10 dB clamp fires."""
    rng = np.random.default_rng(seed)
    spec = _cnoise(rng, n, 1.0) * np.sqrt(n)
    k = np.arange(n)
    lo = (k > frac * n)
    spec[lo] *= 10.0 ** (floor_db / 20.0)
    return np.fft.ifft(spec).astype(np.complex64)


def impulse(n: int, pos: int = 0, amp: complex = 1.0 + 0.5j) -> np.ndarray:
    x = np.zeros(n, dtype=np.complex64)
    x[pos] = amp
    return x


KINDS = {
    "noise": noise,
    "tone40": lambda n, seed=0: tone_noise(n, seed, dbc=-40.0),
    "tone60": lambda n, seed=0: tone_noise(n, seed, dbc=-60.0),
    "wbfm": wbfm,
    "am": am_tone,
    "ssb": ssb_two_tone,
    "halfband": halfband,
}


def make(kind: str, n: int, seed: int = 0, **kw) -> np.ndarray:
    """Dispatch by name; used by tests/golden so inputs never need to be stored."""
    return KINDS[kind](n, seed, **kw) if kw else KINDS[kind](n, seed)
