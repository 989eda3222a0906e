"""Per-call latency of the drop-in module (one SDR read per call, host numpy in / out) next to the
oracle port of the reference functions on one host core.  Prints a small table."""
import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pyspecsdr_b200 import signal_processing as sp, synth
from oracle import ref_dsp as O


def best(fn, reps=30):
    fn(); fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


print(f"{'call':34s} {'ours ms':>9s} {'oracle ms':>10s} {'ratio':>7s}")
for N in (8192, 32768, 262144):
    x = synth.make("wbfm", N, seed=1)
    rows = [("compute_fft", lambda: sp.compute_fft(x), lambda: O.psd_db(x)),
            ("measure_signal_power", lambda: sp.measure_signal_power(x), lambda: O.signal_power_db(x))]
    for mode, fs in (("NFM", 2.4e6), ("WFM", 2.4e6), ("AM", 1e6), ("USB", 1e6)):
        rows.append((f"demodulate_signal {mode}", lambda m=mode, f=fs: sp.demodulate_signal(x, f, m),
                     lambda m=mode, f=fs: O.demod(x, f, m)))
    for name, ours, ref in rows:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a, b = best(ours), best(ref, reps=5)
        print(f"{name + ' N=' + str(N):34s} {a:9.3f} {b:10.3f} {b / a:7.1f}")
