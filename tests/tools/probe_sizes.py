import sys, numpy as np, warnings
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pyspecsdr_b200 import core, synth
from oracle import ref_dsp as O
ctx = core.Context(0)
for s in range(5, 13):
    N = (2 ** s) * 256
    x = synth.make("wbfm", N, seed=s)
    line = f"SAMPLES={s} N={N}:"
    try:
        got = ctx.psd(x, epilogue=True, W=200, want_stats=True)["db"][0]
        want = O.psd_epilogue(O.psd_db(x))
        line += f" psd ok {np.max(np.abs(got-want)):.1e}"
    except Exception as e:
        line += f" psd FAIL({type(e).__name__})"
    for mode, fs in (("NFM", 2.4e6), ("WFM", 2.4e6), ("AM", 1e6), ("USB", 1e6), ("RAW", 2.4e6)):
        try:
            a = ctx.demod(x, fs, mode)[0]
            ref = O.demod(x, fs, mode)
            ref = ref if ref.ndim == 1 else ref[:, 0]
            a0 = a[:, 0] if a.ndim == 2 else a
            line += f" {mode} ok {np.sqrt(np.mean((a0-ref)**2)):.1e}"
        except Exception as e:
            line += f" {mode} FAIL({type(e).__name__}: {str(e)[:40]})"
    print(line, flush=True)
