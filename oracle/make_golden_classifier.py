"""Golden vectors for the signal classifier.  TEST INFRASTRUCTURE ONLY.

`classify_signal` raises NameError in the reference as shipped (`welch` is never imported,
signal_processing.py:299).  This script imports the UNMODIFIED /root/reference/signal_processing.py,
supplies that one missing name (`ref.welch = scipy.signal.welch`) and runs the reference's own
`classify_signal`, `estimate_bandwidth`, `estimate_modulation_index` and the flatness expression (:304,
exec'd from the source line) on seeded inputs; outputs go to tests/golden/classifier.npz.
Run in the build container:  python oracle/make_golden_classifier.py
"""
import os
import sys

import numpy as np
import scipy
import scipy.signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
from pyspecsdr_b200 import synth  # noqa: E402

CASES = [
    # name, generator kwargs, n, fs
    ("wbfm_2400k", dict(kind="wbfm"), 32768, 2.4e6),
    ("wbfm_scan2048", dict(kind="wbfm"), 2048, 2.4e6),
    ("noise_2400k", dict(kind="noise"), 32768, 2.4e6),
    ("tone40_1000k", dict(kind="tone40"), 32768, 1e6),
    ("am_1000k", dict(kind="am"), 32768, 1e6),
    ("ssb_1000k", dict(kind="ssb"), 32768, 1e6),
    ("nfm_250k", dict(kind="wbfm", fs=250e3, dev=5e3, fm=1e3), 32768, 250e3),
    ("nfm_48k", dict(kind="wbfm", fs=48e3, dev=2.5e3, fm=400.0), 16384, 48e3),
    ("am_48k", dict(kind="am", fs=48e3, fm=3e3, depth=0.8), 16384, 48e3),
    ("ssb_24k", dict(kind="ssb", fs=24e3), 8192, 24e3),
    ("halfband_1000k", dict(kind="halfband"), 16384, 1e6),
    ("odd_len_noise", dict(kind="noise"), 5000, 2.4e6),
]


def make_case(kw, n, seed):
    kw = dict(kw)
    kind = kw.pop("kind")
    return synth.make(kind, n, seed=seed, **kw)


def main():
    import signal_processing as ref
    ref.welch = scipy.signal.welch                      # the missing import, nothing else is touched
    src = open(os.path.join(REF, "signal_processing.py")).read().split("\n")
    flat_src = src[303].strip()                         # line 304
    assert flat_src.startswith("spectral_flatness ="), flat_src
    out = {"_numpy": np.array(np.__version__), "_scipy": np.array(scipy.__version__)}
    for i, (name, kw, n, fs) in enumerate(CASES):
        x = make_case(kw, n, seed=20 + i)
        label = ref.classify_signal(x, fs, None)
        freqs, psd = ref.welch(x, fs=fs, nperseg=1024)
        bw = ref.estimate_bandwidth(psd, freqs)
        mi = ref.estimate_modulation_index(x)
        env = {"np": np, "psd": psd}
        exec(flat_src, env)
        out[name + "_label"] = np.array(label)
        out[name + "_feat"] = np.array([bw, mi, env["spectral_flatness"]], dtype=np.float64)
        out[name + "_psd"] = psd.astype(np.float32)
        print(f"{name:18s} {label:13s} bw={bw:12.1f} mi={mi:10.4g} flat={env['spectral_flatness']:.4g}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "classifier.npz"), **out)


if __name__ == "__main__":
    main()
