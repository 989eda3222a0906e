"""CPU ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product package.

A numpy/scipy restatement of the PySpecSDR hot path.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module, and only as the checker / the timed CPU arm.

Where the arithmetic lives: the reference itself contains no numeric kernels; every
number is produced by numpy (pocketfft `np.fft.fft`, `np.hamming`, `np.angle`,
`np.interp`, `np.median`, `np.convolve`) and scipy.signal (`firwin`, `lfilter`,
`butter`, `cheby1`, `sosfilt`, `sosfiltfilt`, `decimate`, `hilbert`), pinned by the
reference only as `numpy>=1.20.0`, `scipy>=1.7.0` (requirements.txt:1-2).  This image
has numpy 2.3.5 / scipy 1.18.1, and the oracle calls the SAME library routines in the
SAME order and dtypes as the reference does (NEP-50 promotion included: a complex64
block stays complex64 through `iq_correction`, the FM discriminator is float32).

Pinning: the reference ships no tests and no golden vectors.  The pin is
`oracle/make_golden.py`, which imports `/root/reference/signal_processing.py` and (with
stubbed SoapySDR/sounddevice/curses) `/root/reference/pyspecsdr.py` in the build
container, runs them on the seeded inputs of `pyspecsdr_b200/synth.py`, and commits the
outputs under `tests/golden/`.  `tests/test_oracle_golden.py` checks every function here
against those files, so the oracle is pinned to outputs of the reference itself.

All citations are into /root/reference/.
"""
from __future__ import annotations

import numpy as np
from scipy import signal as _sig

AUDIO_RATE = 22050      # pyspecconst.py:3  DEFAULT_SAMPLE_RATE
BUTTER_ORDER = 5        # pyspecconst.py:5
WATERFALL_ROWS = 30     # pyspecsdr.py:131  WATERFALL_MAX_LINES
PERSIST_ROWS = 10       # pyspecsdr.py:152  PERSISTENCE_LENGTH
PERSIST_ALPHA = 0.7     # pyspecsdr.py:151
SURFACE_ANGLE = 45      # pyspecsdr.py:155


# --------------------------------------------------------------------------- PSD (a1, a2, a3)
def psd_db(samples: np.ndarray, window: str = "hamming") -> np.ndarray:
    """a1  compute_fft, signal_processing.py:243-264 (window='hamming');
    a18 scanner spectrum, pyspecsdr.py:2542-2543 / 1050-1051 (window='none').
    Accepts [N] or [F, N]; returns float64 of the same shape."""
    x = np.asarray(samples)
    n = x.shape[-1]
    if window == "hamming":
        x = x * np.hamming(n)                       # :246-247, c64*f64 -> c128
    elif window == "hann":
        x = x * np.hanning(n)                       # north_star variant, not in the reference
    elif window != "none":
        raise ValueError(window)
    spec = np.fft.fftshift(np.fft.fft(x, axis=-1), axes=-1)     # :250
    return 10 * np.log10(np.abs(spec) ** 2 + 1e-10)            # :262


def psd_epilogue(row: np.ndarray) -> np.ndarray:
    """a2  main-loop smoothing + noise clamp, pyspecsdr.py:2278-2283. [N] -> [N-4]."""
    k = 5
    sm = np.convolve(row, np.ones(k) / k, mode="valid")         # :2279
    floor = np.median(sm) - 10                                  # :2282
    sm[sm < floor] = floor                                      # :2283
    return sm


def peak_avg(row: np.ndarray):
    """a3  header read-outs, pyspecsdr.py:388-389 (PEAK_POWER feeds the squelch gate :2261)."""
    return float(np.max(row)), float(np.mean(row))


# --------------------------------------------------------------------------- display (a4-a7)
def resample_cols(row: np.ndarray, width: int) -> np.ndarray:
    """The W-column linear resample used by every draw_* (pyspecsdr.py:448-452, 1378-1382,
    1547-1551, 1583-1587, 1682-1684)."""
    n = len(row)
    return np.interp(np.linspace(0, n - 1, width), np.arange(n), row)


def stack_range(rows):
    """Finite min/max over the history stack (pyspecsdr.py:1356-1358, 1525-1527, 1654-1656)."""
    a = np.array(rows)
    fin = a[np.isfinite(a)]
    return float(np.min(fin)), float(np.max(fin))


def waterfall_accumulate(history: list, row: np.ndarray, width: int, max_rows: int = WATERFALL_ROWS):
    """a4  draw_waterfall numeric part, pyspecsdr.py:1351-1358, 1373-1398.
    Mutates `history` like the global list.  Returns (norm[rows, W] newest first, (min, max),
    colour_index[rows, W] int, level[rows, W] int in 0..3 for '.', '-', '=', '#')."""
    history.append(row)
    if len(history) > max_rows:
        history.pop(0)
    lo, hi = stack_range(history)
    norm = np.empty((len(history), width))
    for y, line in enumerate(reversed(history)):                # :1373
        norm[y] = (resample_cols(line, width) - lo) / (hi - lo)  # :1387
    colour = (norm * 5).astype(np.int64)                        # :1388 int() truncation
    level = (norm > 0.25).astype(np.int64) + (norm > 0.5) + (norm > 0.75)   # :1390-1397
    return norm, (lo, hi), colour, level


def gradient_accumulate(history: list, row: np.ndarray, width: int, max_rows: int = WATERFALL_ROWS):
    """a4 (gradient variant)  draw_gradient_waterfall, pyspecsdr.py:1649-1696.
    Returns (norm, (min, max), char_index 0..8, colour_index)."""
    history.append(row)
    if len(history) > max_rows:
        history.pop(0)
    lo, hi = stack_range(history)
    rng = hi - lo
    if rng == 0:
        rng = 1                                                 # :1657-1659
    norm = np.empty((len(history), width))
    for y, line in enumerate(reversed(history)):
        norm[y] = (resample_cols(line, width) - lo) / rng       # :1688
    chars = (norm * 8).astype(np.int64)                         # :1691, 9 glyphs
    colour = (norm * 5).astype(np.int64)                        # :1695
    return norm, (lo, hi), chars, colour


def persistence_accumulate(history: list, row: np.ndarray, width: int, height: int,
                           max_rows: int = PERSIST_ROWS):
    """a5  draw_persistence numeric part, pyspecsdr.py:1521-1556.
    Returns (y[rows, W] int screen rows (oldest trace first), colour_pair[rows] int, (min, max))."""
    history.append(row)
    if len(history) > max_rows:
        history.pop(0)
    lo, hi = stack_range(history)
    rng = hi - lo
    if rng == 0:
        rng = 1                                                 # :1528-1530
    ys = np.empty((len(history), width), dtype=np.int64)
    colours = np.empty(len(history), dtype=np.int64)
    for i, line in enumerate(history):                          # :1543
        alpha = PERSIST_ALPHA ** (max_rows - i)                 # :1544
        colours[i] = int(1 + (5 * (1 - alpha)))                 # :1545
        nrm = (resample_cols(line, width) - lo) / rng           # :1555
        ys[i] = ((1 - nrm) * (height - 1)).astype(np.int64)     # :1556 int() truncation
    return ys, colours, (lo, hi)


def surface_row(row: np.ndarray, width: int):
    """a6  draw_surface_plot numeric part, pyspecsdr.py:1575-1596.
    Returns (magnitude[W] int, (min, max))."""
    fin = row[np.isfinite(row)]
    lo, hi = float(np.min(fin)), float(np.max(fin))
    rng = hi - lo
    if rng == 0:
        rng = 1
    nrm = (row - lo) / rng                                      # :1580
    cols = resample_cols(nrm, width)                            # :1583-1587
    return (cols * 20).astype(np.int64), (lo, hi)              # :1593


def spectrum_normalise(row: np.ndarray, width: int):
    """a7  draw_spectrogram numeric part, pyspecsdr.py:418-452.
    Returns (cols[W] float in [0,1], (display_min, display_max))."""
    fin = row[np.isfinite(row)]
    hi = np.max(fin)
    floor = np.percentile(fin, 20)                              # :422
    span = hi - floor
    dmin = floor - span * 0.1                                   # :426
    dmax = hi + span * 0.05                                     # :427
    nrm = np.clip((row - dmin) / (dmax - dmin), 0, 1)           # :442
    nrm = np.power(nrm, 0.7)                                    # :445
    return resample_cols(nrm, width), (float(dmin), float(dmax))


# --------------------------------------------------------------------------- helpers (a14-a17, a19)
def iq_correct(x: np.ndarray) -> np.ndarray:
    """a14  iq_correction, signal_processing.py:46-80."""
    p_in = np.var(x - np.mean(x))                               # :48-49
    q_amp = np.sqrt(2 * np.mean(x.imag ** 2))                   # :52
    z = x / q_amp                                               # :55
    i, q = z.real, z.imag
    alpha = np.sqrt(2 * np.mean(i ** 2))                        # :60
    sin_phi = (2 / alpha) * np.mean(i * q)                      # :61
    cos_phi = np.sqrt(1 - sin_phi ** 2)                         # :64
    i2 = (1 / alpha) * i                                        # :67
    q2 = (-sin_phi / alpha) * i + q                             # :68
    c = (i2 + 1j * q2) / cos_phi                                # :71
    return c * np.sqrt(p_in / np.var(c))                        # :80


def bandpass(data, lo, hi, fs):
    """a15  bandpass_filter, signal_processing.py:34-42."""
    nyq = fs / 2
    if lo <= 0:
        sos = _sig.butter(BUTTER_ORDER, hi / nyq, btype="low", output="sos")
    else:
        sos = _sig.butter(BUTTER_ORDER, [lo / nyq, hi / nyq], btype="band", output="sos")
    return _sig.sosfilt(sos, data)


def stereo(mono: np.ndarray) -> np.ndarray:
    """a16  mono_to_stereo, signal_processing.py:83-88."""
    out = np.zeros((len(mono), 2))
    out[:, 0] = mono
    out[:, 1] = mono
    return out


def signal_power_db(x: np.ndarray) -> float:
    """a17  measure_signal_power, signal_processing.py:325-328."""
    return 10 * np.log10(np.mean(np.abs(x) ** 2) + 1e-10)


def to_int16(audio: np.ndarray) -> np.ndarray:
    """a19  write_audio_samples numeric line, audio_processing.py:36-38 (C truncation)."""
    return np.int16(audio * 32767)


# --------------------------------------------------------------------------- demodulators (a8-a13)
def _discriminator(x):
    return np.angle(x[1:] * np.conj(x[:-1]))                    # signal_processing.py:94, :122


def demod_nfm(x, fs, audio_rate=AUDIO_RATE):
    """a9  demodulate_nfm, signal_processing.py:91-116."""
    d = _discriminator(x) * (fs / (2 * np.pi))                  # :94-97
    taps = _sig.firwin(numtaps=65, cutoff=15000 / (fs / 2))     # :105-107
    y = _sig.lfilter(taps, 1.0, d)                              # :108
    y = _sig.decimate(y, int(fs / audio_rate))                  # :111-112
    return stereo(y / np.max(np.abs(y)) * 0.95)                 # :115-116


def demod_wfm(x, fs, audio_rate=AUDIO_RATE):
    """a10  demodulate_wfm, signal_processing.py:119-176 (the RDS block :166-174 raises
    NameError inside its own try/except and has no effect)."""
    d = _discriminator(x)                                       # :122
    mono = bandpass(d, 0, 15000, fs)                            # :126
    pilot = bandpass(d, 19000 - 200, 19000 + 200, fs)           # :129
    pilot = np.sin(np.unwrap(np.angle(_sig.lfilter([1], [1, -0.99], pilot))))   # :130
    diff = bandpass(d, 38000 - 15000, 38000 + 15000, fs) * (2 * pilot)          # :133-134
    diff = bandpass(diff, 0, 15000, fs)                         # :137
    left, right = (mono + diff) / 2, (mono - diff) / 2          # :140-141
    a = np.exp(-1 / (75e-6 * fs))                               # :144-145
    left = _sig.lfilter([1 - a], [1, -a], left)                 # :148
    right = _sig.lfilter([1 - a], [1, -a], right)               # :149
    q = int(fs / audio_rate)                                    # :152
    if q > 1:
        left = _sig.decimate(left, q, zero_phase=True)          # :154
        right = _sig.decimate(right, q, zero_phase=True)        # :155
    peak = max(np.max(np.abs(left)), np.max(np.abs(right)))     # :158
    return np.column_stack((left / peak, right / peak))         # :159-163


def demod_am(x):
    """a11  demodulate_am, signal_processing.py:179-195 (filter designed at 22 050 Hz)."""
    env = np.abs(x)                                             # :182
    env = env - np.mean(env)                                    # :185
    y = bandpass(env, 300.0, 3000.0, AUDIO_RATE)                # :188-191
    return stereo(y / np.max(np.abs(y)) * 0.95)                 # :194-195


def demod_ssb(x, fs, lower=True):
    """a12  demodulate_ssb, signal_processing.py:198-217 (both branches are identical)."""
    taps = _sig.firwin(65, 3000 / fs, window="hamming")         # :203 / :208
    z = _sig.lfilter(taps, 1.0, x)                              # :204 / :209
    y = np.real(_sig.hilbert(np.real(z)))                       # :205,:213
    return stereo(y / np.max(np.abs(y)) * 0.95)                 # :216-217


def demod(x, fs, mode="NFM"):
    """a8  demodulate_signal dispatch, signal_processing.py:220-240."""
    if mode not in ("NFM", "AM", "USB", "LSB"):
        x = iq_correct(x)                                       # :222-225
    if mode == "NFM":
        return demod_nfm(x, fs)
    if mode == "WFM":
        return demod_wfm(x, fs)
    if mode == "AM":
        return demod_am(x)
    if mode == "USB":
        return demod_ssb(x, fs, lower=False)
    if mode == "LSB":
        return demod_ssb(x, fs, lower=True)
    if mode == "RAW":
        return np.real(x)                                       # :238
    return np.zeros((len(x), 2))                                # :240


# --------------------------------------------------------------------------- scanner (a18)
def scan_step(samples: np.ndarray, fs: float, threshold=None):
    """a18  per-step scanner numerics.
    threshold=None: inline 'c'-key loop, pyspecsdr.py:2542-2552 (mask = dB > peak-20);
    otherwise scan_frequencies, pyspecsdr.py:1050-1057 (mask = dB > threshold).
    Returns (peak_db, n_above, bandwidth_hz)."""
    db = psd_db(samples, window="none")
    peak = np.max(db)
    mask = db > (peak - 20) if threshold is None else db > threshold
    count = int(np.sum(mask))
    return float(peak), count, count * (fs / len(db))


# --------------------------------------------------------------------------- classifier (§8f-4)
def classify_features(samples: np.ndarray, sample_rate: float):
    """classify_signal's three features, signal_processing.py:267-304, with the one name the
    reference forgets to import (`welch`, :299) supplied from scipy.signal.  Returns
    (signal_bw, modulation_index, spectral_flatness) in the reference's dtypes."""
    freqs, psd = _sig.welch(samples, fs=sample_rate, nperseg=1024)                       # :299
    psd_db = 10 * np.log10(psd + 1e-10)                                                  # :270
    mask = psd_db > (np.max(psd_db) + (-20))                                             # :271-274
    fr = freqs[mask]
    signal_bw = fr[-1] - fr[0] if np.any(mask) else 0                                    # :275-280
    amp_var = np.var(np.abs(samples))                                                    # :286,290
    phase_var = np.var(np.diff(np.unwrap(np.angle(samples))))                            # :287,291
    modulation_index = phase_var / (amp_var + 1e-10)                                     # :293
    spectral_flatness = np.exp(np.mean(np.log(psd + 1e-10))) / np.mean(psd)              # :304
    return signal_bw, modulation_index, spectral_flatness


def classify_label(signal_bw, modulation_index, spectral_flatness) -> str:
    """The decision tree of classify_signal, signal_processing.py:306-322."""
    if signal_bw > 150e3:
        if modulation_index > 0.8:
            return 'FM_BROADCAST'
    elif 8e3 <= signal_bw <= 16e3:
        if modulation_index < 0.3:
            return 'NARROW_FM'
    elif 8e3 <= signal_bw <= 10e3:
        if modulation_index < 0.2 and spectral_flatness < 0.3:
            return 'AM_BROADCAST'
    elif 2e3 <= signal_bw <= 3e3:
        if spectral_flatness < 0.2:
            return 'SSB'
    elif spectral_flatness > 0.7:
        return 'DIGITAL'
    return 'UNKNOWN'


def classify_signal(samples: np.ndarray, sample_rate: float, bandwidth=None) -> str:
    return classify_label(*classify_features(samples, sample_rate))
