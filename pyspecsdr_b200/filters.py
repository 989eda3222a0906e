"""Host-side filter design and table construction for the demodulation kernels.

Filter *design* stays on the host and uses scipy itself, so coefficients are bit-identical to the
reference's (`firwin`, `butter`, `cheby1`, `sosfilt_zi`; signal_processing.py:34-42, 107, 144-149,
203 and scipy.signal.decimate/sosfiltfilt as called from :112, :154-155).  They depend only on
(mode, sample_rate, block length) and are cached by the caller.

What this module adds is the *chunk-table* form of the decimating chain that the CUDA kernel runs
(DESIGN.md "Demodulation, decimating modes").  For NFM/WFM only every q-th sample of a zero-phase
(forward+backward) 8th-order Chebyshev filter is kept, so instead of running 2x4 biquad recurrences
over every sample the kernel evaluates, per chunk of q samples, a small set of fp64 dot products
against precomputed response tables and then propagates three tiny state recurrences over the
chunk sequence.  Every table is obtained by *probing*: the exact scipy procedure (odd extension of
27 samples, `sosfilt_zi` initial conditions scaled by the first sample, forward pass, reversed
pass) is executed on unit inputs / unit states, so the tables are the reference computation's own
linear map — no re-derivation of filter algebra.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy import signal as sig

AUDIO_RATE = 22050      # pyspecconst.py:3
BUTTER_ORDER = 5        # pyspecconst.py:5
FIR_TAPS = 65           # signal_processing.py:107, :203
EDGE = 27               # sosfiltfilt default padlen for 4 sections: 3 * (2*4 + 1)

MODES = {"NFM": 0, "WFM": 1, "AM": 2, "USB": 3, "LSB": 4, "RAW": 5}


def decimation_factor(fs: float) -> int:
    return int(fs / AUDIO_RATE)                      # signal_processing.py:111, :152


def nfm_taps(fs: float) -> np.ndarray:
    return sig.firwin(numtaps=FIR_TAPS, cutoff=15000 / (fs / 2))      # :105-107


def ssb_taps(fs: float) -> np.ndarray:
    return sig.firwin(FIR_TAPS, 3000 / fs, window="hamming")          # :203 / :208


def butter_sos(lo: float, hi: float, fs: float) -> np.ndarray:
    nyq = fs / 2                                                       # :34-42
    if lo <= 0:
        return sig.butter(BUTTER_ORDER, hi / nyq, btype="low", output="sos")
    return sig.butter(BUTTER_ORDER, [lo / nyq, hi / nyq], btype="band", output="sos")


def am_sos() -> np.ndarray:
    return butter_sos(300.0, 3000.0, AUDIO_RATE)                      # :188-191 (fs fixed at 22050)


def wfm_pre_sos(fs: float) -> np.ndarray:
    """L+R low-pass (:126), the /2 of :140-141 (L-R is identically ~0) and the 75 us
    de-emphasis one-pole (:144-149) as one cascade of 4 sections."""
    lp = butter_sos(0, 15000, fs)
    # the reference also designs the pilot (:129) and L-R (:133) band-passes; their outputs are
    # multiplied by ~0, but the designs raise ValueError when 53 kHz is not below Nyquist, and so do we
    butter_sos(19000 - 200, 19000 + 200, fs)
    butter_sos(38000 - 15000, 38000 + 15000, fs)
    a = math.exp(-1 / (75e-6 * fs))
    de = np.array([[0.5 * (1 - a), 0.0, 0.0, 1.0, -a, 0.0]])
    return np.vstack([lp, de])


def decim_sos(q: int) -> np.ndarray:
    return sig.cheby1(8, 0.05, 0.8 / q, output="sos")                 # scipy decimate(): ftype='iir', n=8


# ------------------------------------------------------------------------------------------------
@dataclass
class DecimPlan:
    """Everything the decimating-demod kernel needs for one (mode, fs, N)."""
    mode: str
    fs: float
    N: int                 # IQ samples per block
    L: int                 # discriminator samples = N - 1
    q: int
    n_out: int             # ceil(L / q) audio frames per block
    lead: int              # FIR history the body window needs (64 for NFM, 0 for WFM)
    SF: int                # forward-flowing state size (pre-filter states + Chebyshev forward states)
    SB: int                # backward state size (8)
    n_body: int            # body chunks 1..n_body, each produces output k = j
    m_tail: int            # outputs produced inside the tail block
    tail_start: int        # first discriminator index the tail's window needs minus lead (p index)
    tail_len: int          # tail window length (discriminator samples incl. lead)
    scale: float           # discriminator scale (fs/2pi for NFM, 1 for WFM), applied in fp32
    norm: float            # 0.95 (NFM) or 1.0 (WFM)
    body: np.ndarray = field(repr=False, default=None)    # [(SF+SB+1), q+lead]: rows TF | TB | TR
    AF: np.ndarray = field(repr=False, default=None)      # [SF, SF]
    AB: np.ndarray = field(repr=False, default=None)      # [SB, SB]
    MB: np.ndarray = field(repr=False, default=None)      # [SB, SF]
    CR: np.ndarray = field(repr=False, default=None)      # [SF]
    CB: np.ndarray = field(repr=False, default=None)      # [SB]
    DB: float = 0.0
    head: np.ndarray = field(repr=False, default=None)    # [(SF+1), 28] applied to d[0..27]
    tail_T: np.ndarray = field(repr=False, default=None)  # [(SB+m_tail), tail_len]
    tail_M: np.ndarray = field(repr=False, default=None)  # [(SB+m_tail), SF]


def _sosfilt_state(sos, x, zi):
    y, zf = sig.sosfilt(sos, x, zi=zi.reshape(-1, 2))
    return y, zf.reshape(-1)


def build_decim_plan(mode: str, fs: float, N: int) -> DecimPlan:
    assert mode in ("NFM", "WFM")
    q = decimation_factor(fs)
    L = N - 1
    if q < 2:
        raise ValueError("decimating demodulators need sample_rate >= 2 * 22050")
    if L <= EDGE:
        raise ValueError("block too short for the zero-phase decimator (needs N > 28)")
    n_out = -(-L // q)
    J = n_out - 1                                   # outputs k = 0..J at p-index k*q
    sos_ch = decim_sos(q)
    zi_ch = sig.sosfilt_zi(sos_ch).reshape(-1)      # (4,2) -> 8
    SB = 8
    if mode == "NFM":
        taps = nfm_taps(fs)
        lead, sos_pre, npre = FIR_TAPS - 1, None, 0
        scale, norm = fs / (2 * np.pi), 0.95
    else:
        taps = None
        sos_pre = wfm_pre_sos(fs)
        lead, npre = 0, 2 * len(sos_pre)
        scale, norm = 1.0, 1.0
    SF = npre + 8

    def prefilter(dwin, pre_state):
        """dwin includes `lead` history samples; returns p for the non-history part, new pre state."""
        if taps is not None:
            full = np.convolve(dwin, taps)          # full[i] = sum_k taps[k] dwin[i-k]
            return full[lead:len(dwin)], pre_state
        return _sosfilt_state(sos_pre, dwin, pre_state)

    # tail: the last m_tail chunks plus the remainder and the 27-sample odd extension, merged so
    # that the extension only needs p-samples inside the block
    m_tail = 1
    while m_tail <= J and (L - 1 - (J - m_tail) * q) < EDGE + 1:
        m_tail += 1
    n_body = J - m_tail
    if n_body < 0:
        raise ValueError("block too short for this sample rate (needs more than ~2 decimated frames)")
    tail_p0 = n_body * q + 1                        # first p index inside the tail
    tail_np = L - tail_p0                           # p samples in the tail
    assert tail_np >= EDGE + 1
    tail_len = tail_np + lead

    # ---------------------------------------------------------------- body chunk by probing
    def body_map(s0, t0, dwin):
        p, pre1 = prefilter(dwin, s0[:npre])
        yf, ch1 = _sosfilt_state(sos_ch, p, s0[npre:])
        ybr, t1 = _sosfilt_state(sos_ch, yf[::-1], t0)
        return np.concatenate([pre1, ch1]), t1, yf[-1], ybr[0]

    wlen = q + lead
    nin = SF + SB + wlen
    resp = np.zeros((SF + SB + 2, nin))
    for c in range(nin):
        e = np.zeros(nin)
        e[c] = 1.0
        s1, t1, yl, yb = body_map(e[:SF].copy(), e[SF:SF + SB].copy(), e[SF + SB:].copy())
        resp[:, c] = np.concatenate([s1, t1, [yl, yb]])
    AF = resp[:SF, :SF]
    TF = resp[:SF, SF + SB:]
    MB = resp[SF:SF + SB, :SF]
    AB = resp[SF:SF + SB, SF:SF + SB]
    TB = resp[SF:SF + SB, SF + SB:]
    CR = resp[SF + SB, :SF]
    TR = resp[SF + SB, SF + SB:]
    CB = resp[SF + SB + 1, SF:SF + SB]
    DB = float(np.prod(sos_ch[:, 0]))               # y = C t + D x with D = product of b0
    body = np.vstack([TF, TB, TR[None, :]])

    # ---------------------------------------------------------------- head by probing
    def head_map(d28):
        if taps is not None:
            p = np.convolve(d28, taps)[:EDGE + 1]
            pre_after0 = np.zeros(0)
        else:
            p, _ = _sosfilt_state(sos_pre, d28, np.zeros(npre))
            _, pre_after0 = _sosfilt_state(sos_pre, d28[:1], np.zeros(npre))
        ext = np.concatenate([2 * p[0] - p[EDGE:0:-1], p[:1]])      # ext[0..27]
        yf, ch = _sosfilt_state(sos_ch, ext, zi_ch * ext[0])
        return np.concatenate([pre_after0, ch, [yf[-1]]])
    head = np.zeros((SF + 1, EDGE + 1))
    for c in range(EDGE + 1):
        e = np.zeros(EDGE + 1)
        e[c] = 1.0
        head[:, c] = head_map(e)

    # ---------------------------------------------------------------- tail by probing
    out_pos = [(n_body + 1 + i) * q - tail_p0 for i in range(m_tail)]   # positions inside the tail

    def tail_map(s0, dwin):
        p, _ = prefilter(dwin, s0[:npre])
        ext = np.concatenate([p, 2 * p[-1] - p[-2:-EDGE - 2:-1]])
        yf, _ = _sosfilt_state(sos_ch, ext, s0[npre:])
        ybr, t1 = _sosfilt_state(sos_ch, yf[::-1], zi_ch * yf[-1])
        yb = ybr[::-1]
        return np.concatenate([t1, yb[out_pos]])
    tail_T = np.zeros((SB + m_tail, tail_len))
    tail_M = np.zeros((SB + m_tail, SF))
    for c in range(SF):
        e = np.zeros(SF)
        e[c] = 1.0
        tail_M[:, c] = tail_map(e, np.zeros(tail_len))
    for c in range(tail_len):
        e = np.zeros(tail_len)
        e[c] = 1.0
        tail_T[:, c] = tail_map(np.zeros(SF), e)

    return DecimPlan(mode=mode, fs=fs, N=N, L=L, q=q, n_out=n_out, lead=lead, SF=SF, SB=SB,
                     n_body=n_body, m_tail=m_tail, tail_start=tail_p0 - lead, tail_len=tail_len,
                     scale=scale, norm=norm, body=body, AF=AF, AB=AB, MB=MB, CR=CR, CB=CB, DB=DB,
                     head=head, tail_T=tail_T, tail_M=tail_M)


def matrix_power_seq(A: np.ndarray, n: int) -> np.ndarray:
    """A^n by n-1 successive products (the same rounding pattern as stepping the state n times)."""
    P = np.array(A, dtype=np.float64)
    for _ in range(n - 1):
        P = A @ P
    return P


def emulate_decim(plan: DecimPlan, d: np.ndarray) -> np.ndarray:
    """numpy restatement of what the CUDA kernel does with a plan (fp64), for CPU-side validation of
    the tables: d = scaled discriminator samples [L] -> un-normalised decimated audio [n_out]."""
    q, lead, SF, SB = plan.q, plan.lead, plan.SF, plan.SB
    d = np.asarray(d, dtype=np.float64)
    dpad = np.concatenate([np.zeros(lead), d])               # dpad[i + lead] = d[i]
    hv = plan.head @ d[:EDGE + 1]
    s = [None] * (plan.n_body + 2)
    s[1] = hv[:SF]
    yf27 = hv[SF]
    u = np.zeros((plan.n_body + 1, SF + SB + 1))
    for j in range(1, plan.n_body + 1):
        p0 = (j - 1) * q + 1                                  # first p index of chunk j
        u[j] = plan.body @ dpad[p0:p0 + q + lead]
        s[j + 1] = plan.AF @ s[j] + u[j, :SF]
    ts = plan.tail_start + lead
    tv = plan.tail_M @ s[plan.n_body + 1] + plan.tail_T @ dpad[ts:ts + plan.tail_len]
    t = tv[:SB]
    y = np.zeros(plan.n_out)
    y[plan.n_body + 1:] = tv[SB:]
    for j in range(plan.n_body, 0, -1):
        yf_last = plan.CR @ s[j] + u[j, SF + SB]
        y[j] = plan.CB @ t + plan.DB * yf_last
        t = plan.AB @ t + plan.MB @ s[j] + u[j, SF:SF + SB]
    y[0] = plan.CB @ t + plan.DB * yf27
    return y
