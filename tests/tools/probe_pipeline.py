import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pyspecsdr_b200 import core, synth
from oracle import ref_dsp as O
ctx = core.Context(0)
blocks = np.stack([synth.make("wbfm", 32768, seed=s) for s in range(6)])
for n_fft in (4096, 8192, 32768):
    for mode in ("WFM", "NFM", "AM"):
        try:
            fs = 2.4e6 if mode != "AM" else 1e6
            out = ctx.pipeline(blocks, fs, mode, n_fft, 200, 30)
            ref = O.demod(blocks[1], fs, mode)
            a = out["audio"][1]
            r = ref if ref.ndim == 1 else ref[:, :a.shape[1]] if a.ndim == 2 else ref[:, 0]
            err = float(np.sqrt(np.mean((a.reshape(r.shape) - r) ** 2))) if a.size == r.size else -1
            want = O.psd_epilogue(O.psd_db(blocks[1][:n_fft]))
            fpb = 32768 // n_fft
            cols = out["cols"][1 * fpb]
            cerr = float(np.max(np.abs(cols - O.resample_cols(want, 200))))
            print(n_fft, mode, "ok audio", a.shape, f"{err:.1e}", "cols", f"{cerr:.1e}", {k: v.shape for k, v in out.items()})
        except Exception as e:
            print(n_fft, mode, "FAIL", type(e).__name__, str(e)[:100])
