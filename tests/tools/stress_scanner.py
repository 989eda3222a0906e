"""Randomised stress of the scanner's integer bin count (outside pytest): random signal kinds, sizes 512 ... 8192,
the inline 'peak - 20 dB' rule and absolute thresholds; the count must EQUAL the oracle's (fp64 dB comparison in the
reference, fp64 power comparison here: they differ only on ties at the 1e-15 level), the peak within 1e-4 dB.
usage: python tests/tools/stress_scanner.py [frames_per_case]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_dsp as O                      # noqa: E402
from pyspecsdr_b200 import core, synth               # noqa: E402


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    ctx = core.Context(0)
    rng = np.random.default_rng(5)
    total = bad = 0
    for n in (512, 1024, 2048, 4096, 8192):
        for kind in ("noise", "tone40", "tone60", "wbfm", "halfband"):
            x = np.stack([synth.make(kind, n, seed=int(rng.integers(1 << 30))) * np.float32(10.0 ** rng.uniform(-2, 2))
                          for _ in range(per)]).astype(np.complex64)
            peak, count = ctx.scan(x)
            thr = float(rng.uniform(-20, 40))
            peak_a, count_a = ctx.scan(x, threshold=thr)
            for f in range(per):
                pk, c, _ = O.scan_step(x[f], 2.4e6)
                _, ca, _ = O.scan_step(x[f], 2.4e6, threshold=thr)
                assert abs(peak[f] - pk) <= 1e-4 and abs(peak_a[f] - pk) <= 1e-4, (n, kind, f)
                bad += int(count[f] != c) + int(count_a[f] != ca)
                total += 2
    print(f"stress {'ok' if bad == 0 else 'FAIL'}: {total} counts, {bad} differ")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
