"""Randomised stress of the smoothing / median-clamp epilogue against the oracle (outside pytest: ~1 min on a B200).
Rows with designed spectra (un-windowed so the design survives) from several distribution families, many seeds,
sizes 512 ... 32768: every output within 1e-4 dB.  usage: python tests/tools/stress_epilogue.py [rows_per_case]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_dsp as O                      # noqa: E402
from pyspecsdr_b200 import core                      # noqa: E402

TOL = 1e-4


def design(n, family, rng):
    k = np.arange(n)
    if family == "gauss":
        db = rng.normal(rng.uniform(-80, 20), rng.uniform(0.01, 12), n)
    elif family == "bimodal":
        db = np.where(rng.random(n) < rng.uniform(0.05, 0.95), rng.uniform(-90, -30), rng.uniform(-30, 40)) + rng.normal(0, rng.uniform(0.01, 3), n)
    elif family == "plateaus":
        steps = rng.integers(2, 40)
        db = -20.0 - 3.0 * np.floor(steps * k / n) + rng.normal(0, 10.0 ** rng.uniform(-6, -1), n)
    elif family == "outliers":
        db = rng.normal(-50, rng.uniform(0.01, 1.0), n)
        m = rng.integers(1, max(2, n // 50))
        db[rng.choice(n, m, replace=False)] = rng.uniform(-99, 60, m)
    elif family == "notch":
        db = rng.normal(-30, 1.5, n)
        a = rng.integers(0, n - n // 4)
        db[a:a + rng.integers(8, n // 4)] = rng.uniform(-99, -60)
    elif family == "ramp":
        db = np.linspace(rng.uniform(-90, -40), rng.uniform(-30, 30), n) + rng.normal(0, rng.uniform(0, 0.5), n)
    else:
        raise ValueError(family)
    spec = 10.0 ** (np.clip(db, -99.0, 60.0) / 20.0) * np.exp(2j * np.pi * rng.random(n))
    return np.fft.ifft(np.fft.ifftshift(spec)).astype(np.complex64)


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    ctx = core.Context(0)
    rng = np.random.default_rng(2026)
    worst, cases = 0.0, 0
    for n in (512, 1024, 2048, 4096, 8192, 16384, 32768):
        for family in ("gauss", "bimodal", "plateaus", "outliers", "notch", "ramp"):
            x = np.stack([design(n, family, rng) for _ in range(rows)])
            for window in ("none", "hamming"):
                res = ctx.psd(x, window=window, epilogue=True, W=157, want_stats=True)
                for f in range(rows):
                    want = O.psd_epilogue(O.psd_db(x[f], window=window))
                    e = float(np.max(np.abs(res["db"][f] - want)))
                    e = max(e, float(np.max(np.abs(res["cols"][f] - O.resample_cols(want, 157)))))
                    pk, av = O.peak_avg(want)
                    e = max(e, abs(res["stats"][f][0] - pk), abs(res["stats"][f][1] - av))
                    if not e <= TOL:
                        print(f"FAIL n={n} family={family} window={window} row={f}: {e:.3e}")
                        sys.exit(1)
                    worst = max(worst, e)
                    cases += 1
    print(f"stress ok: {cases} rows, worst error {worst:.2e} dB (bar {TOL:g})")


if __name__ == "__main__":
    main()
