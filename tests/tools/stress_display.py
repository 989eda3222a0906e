"""Randomised stress of the display accumulate's integer planes against the oracle (outside pytest).
Random fp64 dB rows of random length fed one by one into carried display streams of random geometry; every
normalised value, range, glyph / colour / screen-row plane must be EQUAL to what the oracle (pinned to the
cells the reference drew) computes.  usage: python tests/tools/stress_display.py [trials]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_dsp as O                      # noqa: E402
from pyspecsdr_b200 import core                      # noqa: E402


def rows_for(rng, n, count):
    base, spread = rng.uniform(-90, 0), 10.0 ** rng.uniform(-3, 1.3)
    out = []
    for s in range(count):
        r = rng.normal(base, spread, n)
        if rng.random() < 0.3:
            r[rng.integers(0, n)] += rng.uniform(10, 80)           # a carrier
        if s and rng.random() < 0.15:
            r = out[-1].copy()                                     # a repeated row
        if rng.random() < 0.1 and spread > 1.0:
            r = np.round(r, 1)                                     # many ties, values on round numbers (never a
                                                                   # constant stack: the reference's int(nan) raises there)
        out.append(r)
    return out


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    ctx = core.Context(0)
    rng = np.random.default_rng(77)
    cells = 0
    for t in range(trials):
        n = int(rng.choice([60, 252, 508, 1020, 4092, 8188]))
        W = int(rng.integers(2, 400))
        H = int(rng.integers(4, 60))
        count = int(rng.integers(3, 45))
        rows = rows_for(rng, n, count)
        # waterfall / gradient (30-row history)
        for sid, kind in ((1, "waterfall"), (2, "gradient")):
            ctx.display_open(sid, kind, W=W, rows_max=30)
            hist = []
            for r in rows:
                res = ctx.display_accumulate(sid, r)
                if kind == "waterfall":
                    norm, (lo, hi), b, a = O.waterfall_accumulate(hist, r, W)
                else:
                    norm, (lo, hi), a, b = O.gradient_accumulate(hist, r, W)
                m = len(hist)
                np.testing.assert_array_equal(res["norm64"][0, :m], norm)
                np.testing.assert_array_equal(res["minmax64"][0], [lo, hi])
                np.testing.assert_array_equal(res["plane_a"][0, :m], a)
                np.testing.assert_array_equal(res["plane_b"][0, :m], b)
                assert np.all(res["plane_a"][0, m:] == 255)
                cells += m * W
            ctx.display_close(sid)
        # persistence (10 traces)
        ctx.display_open(3, "persistence", W=W, rows_max=10, H=H)
        hist = []
        for r in rows:
            res = ctx.display_accumulate(3, r)
            ys, colours, (lo, hi) = O.persistence_accumulate(hist, r, W, H)
            m = len(hist)
            np.testing.assert_array_equal(res["plane_a"][0, :m][::-1].astype(np.int64), ys)
            np.testing.assert_array_equal(res["plane_b"][0, :m, 0][::-1].astype(np.int64), colours)
            np.testing.assert_array_equal(res["minmax64"][0], [lo, hi])
            cells += m * W
        ctx.display_close(3)
        # surface (one row)
        ctx.display_open(4, "surface", W=W, rows_max=1)
        res = ctx.display_accumulate(4, np.array(rows))
        for s, r in enumerate(rows):
            mag, (lo, hi) = O.surface_row(r, W)
            np.testing.assert_array_equal(res["plane_a"][s, 0], mag)
            np.testing.assert_array_equal(res["minmax64"][s], [lo, hi])
            cells += W
        ctx.display_close(4)
    print(f"stress ok: {trials} trials, {cells} cells, every plane equal")


if __name__ == "__main__":
    main()
