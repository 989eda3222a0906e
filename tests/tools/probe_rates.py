"""Every (sample rate, block size, mode) combination a user can plausibly select, against the oracle."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pyspecsdr_b200 import core, synth
from oracle import ref_dsp as O
ctx = core.Context(0)
bad = 0
for fs in [float(a) for a in os.environ.get("PROBE_FS", "48e3,250e3,900001,1.024e6,2.048e6,2.4e6,3.2e6,8e6,10e6,20e6,40e6,56e6,61.44e6").split(",")]:
    for N in (8192, 32768, 262144):
        x = synth.wbfm(N, seed=3, fs=fs, dev=min(75e3, fs / 8))
        line = f"fs={fs:>10.0f} N={N:>7d}:"
        for mode in ("NFM", "WFM", "AM", "USB", "RAW"):
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ref = O.demod(x, fs, mode)
            except Exception as e:
                ref = e
            try:
                a = ctx.demod(x, fs, mode)[0]
            except Exception as e:
                a = e
            if isinstance(ref, Exception) or isinstance(a, Exception):
                ok = isinstance(ref, Exception) and isinstance(a, Exception)
                line += f" {mode} {'both-raise' if ok else 'MISMATCH ref=' + type(ref).__name__ + ' ours=' + (type(a).__name__ + ':' + str(a)[:50] if isinstance(a, Exception) else 'ok')}"
                bad += not ok
                continue
            r = ref if ref.ndim == 1 else ref[:, 0]
            g = a[:, 0] if a.ndim == 2 else a
            if r.shape != g.shape:
                line += f" {mode} SHAPE {r.shape} vs {g.shape}"
                bad += 1
                continue
            rms = float(np.sqrt(np.mean((g - r) ** 2)))
            line += f" {mode} {rms:.1e}"
            bad += not (rms <= 1e-5)
    # one line per (fs, N)
        print(line, flush=True)
print("BAD", bad)
