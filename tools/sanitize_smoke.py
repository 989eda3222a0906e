"""Small workload touching every kernel family once (for compute-sanitizer memcheck / racecheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyspecsdr_b200 import core, synth
ctx = core.Context(0)
for n in (512, 1024, 4096, 8192):
    x = np.stack([synth.make("tone40", n, seed=s) for s in range(5)])
    ctx.psd(x)
    ctx.psd(x, epilogue=True, W=200, want_stats=True)
    ctx.psd(np.zeros((2, n), np.complex64), epilogue=True, W=50, want_stats=True)      # flat rows: rare median path
    if n >= 512:
        ctx.scan(x, rel_db=20.0)
for n in (16384, 32768, 65536, 131072, 262144):
    x = np.stack([synth.make("wbfm", n, seed=s) for s in range(2)])
    ctx.psd(x)
    ctx.psd(x, epilogue=True, W=200, want_stats=True)
blk = np.stack([synth.make("wbfm", 32768, seed=s) for s in range(3)])
for mode, fs in (("NFM", 2.4e6), ("WFM", 2.4e6), ("AM", 1e6), ("USB", 1e6), ("RAW", 2.4e6), ("NFM", 250e3), ("WFM", 56e6)):
    ctx.demod(blk, fs, mode)
long = np.stack([synth.make("am", 100000, seed=s) for s in range(2)])
ctx.demod(long, 1e6, "AM")
ctx.demod(long, 1e6, "USB")
ctx.demod(np.stack([synth.make("wbfm", 262144, seed=1)]), 250e3, "NFM")
ctx.classify(blk, 2.4e6)
ctx.signal_power(blk)
ctx.iq_correct(blk[0])
ctx.to_int16(np.zeros((100, 2), np.float32))
res = ctx.psd(blk.reshape(-1, 4096), epilogue=True, W=200, want_stats=True)
ctx.display_render(res["cols"], res["stats"], rows_max=30)
# round 2: stateful display streams (fp64 host rows and float32 device-layout rows), the pipeline with a carried
# ring and planes, exact int16 / spectrum paths, the scan kernel with its scratch in global memory (long block)
rows = res["db"].astype(np.float64)
for sid, (kind, R, H) in enumerate((("waterfall", 30, 0), ("gradient", 30, 0), ("persistence", 10, 36), ("surface", 1, 0))):
    ctx.display_open(sid, kind, W=120, rows_max=R, H=H)
    ctx.display_accumulate(sid, rows[:7])
    ctx.display_accumulate(sid, rows[7:9])
    ctx.display_close(sid)
ctx.display_open(9, "waterfall", W=200, rows_max=30)
ctx.pipeline(blk, 2.4e6, "WFM", n_fft=4096, W=200, rows_max=30, display_stream=9, want_planes=True)
ctx.pipeline(blk[:1], 2.4e6, "NFM", n_fft=4096, W=200, rows_max=30, display_stream=9, want_planes=True)
ctx.display_close(9)
ctx.spectrum_normalise(rows[:3], 113)
ctx.to_int16(np.linspace(-1, 1, 1001))
ctx.demod(blk, 2.4e6, "WFM", iq_correct=False)
ctx.demod(np.stack([synth.make("wbfm", 262144, seed=2)]), 250e3, "WFM")      # 23828 chunks: scratch stays in global memory
ctx.demod(np.stack([synth.make("wbfm", 65536, seed=3)]), 20e6, "NFM")       # q = 907: table read from global memory
print("sanitize workload done, launches", ctx.launches)
