"""GPU parity: display accumulate (waterfall / gradient / persistence history normalisation) vs the
oracle, which is itself pinned through what the reference's draw_* functions drew."""
import numpy as np
import pytest

from oracle import ref_dsp as O
from pyspecsdr_b200 import synth

pytestmark = pytest.mark.gpu


def make_rows(count=34, n=4096):
    x = np.stack([synth.make("wbfm" if s % 2 else "tone40", n, seed=100 + s) for s in range(count)])
    ref_rows = [O.psd_epilogue(O.psd_db(r)) for r in x]
    return x, ref_rows


def test_waterfall_history_vs_oracle(ctx):
    x, ref_rows = make_rows()
    W = 112
    res = ctx.psd(x, epilogue=True, W=W, want_stats=True, want_db=False)
    norm, mm = ctx.display_render(res["cols"], res["stats"], rows_max=30)
    hist = []
    mism = 0
    for s, r in enumerate(ref_rows):
        want, (lo, hi), colour, level = O.waterfall_accumulate(hist, r, W)
        got = norm[s, :len(hist)]
        assert abs(mm[s, 0] - lo) <= 1e-4 and abs(mm[s, 1] - hi) <= 1e-4
        assert np.max(np.abs(got - want)) <= 1e-5
        assert np.all(np.isnan(norm[s, len(hist):]))
        # quantised planes (colour index, glyph level) agree except where norm sits on a boundary
        mism += np.sum((got * 5).astype(np.int64) != colour)
        mism += np.sum(((got > 0.25).astype(int) + (got > 0.5) + (got > 0.75)) != level)
    assert mism <= 1e-4 * 2 * 34 * 30 * W


def test_persistence_history_vs_oracle(ctx):
    x, ref_rows = make_rows(14)
    W, H = 112, 36
    res = ctx.psd(x, epilogue=True, W=W, want_stats=True, want_db=False)
    norm, mm = ctx.display_render(res["cols"], res["stats"], rows_max=10, guard_zero_range=True)
    hist = []
    bad = 0
    for s, r in enumerate(ref_rows):
        ys, colours, (lo, hi) = O.persistence_accumulate(hist, r, W, H)
        got = norm[s, :len(hist)][::-1]                      # oldest first, like PERSISTENCE_HISTORY
        y = ((1 - got.astype(np.float64)) * (H - 1)).astype(np.int64)
        bad += np.sum(y != ys)
        assert np.max(np.abs(y - ys)) <= 1
    assert bad <= 1e-3 * 14 * 10 * W


def test_display_strided_renders(ctx):
    x, ref_rows = make_rows(20, 1024)
    W = 64
    res = ctx.psd(x, epilogue=True, W=W, want_stats=True, want_db=False)
    full, _ = ctx.display_render(res["cols"], res["stats"], rows_max=30)
    some, _ = ctx.display_render(res["cols"], res["stats"], rows_max=30, first=3, step=8, n_renders=3)
    np.testing.assert_array_equal(some, full[[3, 11, 19]])


def test_pipeline_matches_separate_calls_across_chunks(ctx):
    """pss_pipeline_c64 (chunked, overlapped copies) == the individual host-pointer calls, bitwise,
    with more blocks than one copy chunk so the waterfall history crosses chunk boundaries."""
    n_block, n_fft, W = 4096, 1024, 64
    base = np.stack([synth.make("wbfm", n_block, seed=s, fs=1.024e6) for s in range(12)])
    blocks = np.ascontiguousarray(np.tile(base, (50, 1))[:590])            # 590 blocks > 2 chunks of 256
    out = ctx.pipeline(blocks, 1.024e6, "NFM", n_fft=n_fft, W=W, rows_max=30, want_db=True)
    frames = blocks.reshape(-1, n_fft)
    sep = ctx.psd(frames, epilogue=True, W=W, want_stats=True)
    np.testing.assert_array_equal(out["db"], sep["db"])
    np.testing.assert_array_equal(out["cols"], sep["cols"])
    np.testing.assert_array_equal(out["stats"], sep["stats"])
    fpb = n_block // n_fft
    norm, mm = ctx.display_render(sep["cols"], sep["stats"], rows_max=30, first=fpb - 1, step=fpb, n_renders=len(blocks))
    np.testing.assert_array_equal(out["norm"], norm)
    np.testing.assert_array_equal(out["minmax"], mm)
    np.testing.assert_array_equal(out["audio"], ctx.demod(blocks, 1.024e6, "NFM"))


def test_spectrum_normalise_vs_oracle(ctx):
    """a7: 20th-percentile floor, clip, ** 0.7, resample (pyspecsdr.py:418-452)."""
    x, ref_rows = make_rows(6, 4096)
    W = 113
    db = ctx.psd(x, epilogue=True)["db"]
    cols, rng = ctx.spectrum_normalise(db, W)
    for f, r in enumerate(ref_rows):
        want, (dmin, dmax) = O.spectrum_normalise(r, W)
        assert abs(rng[f, 0] - dmin) <= 2e-4 and abs(rng[f, 1] - dmax) <= 2e-4
        assert np.max(np.abs(cols[f] - want)) <= 2e-5
    # long rows (the app's default 32768-sample read -> 32764 bins)
    xl = np.stack([synth.make("wbfm", 32768, seed=s) for s in range(2)])
    dbl = ctx.psd(xl, epilogue=True)["db"]
    cols, rng = ctx.spectrum_normalise(dbl, 200)
    for f in range(2):
        want, (dmin, dmax) = O.spectrum_normalise(O.psd_epilogue(O.psd_db(xl[f])), 200)
        assert np.max(np.abs(cols[f] - want)) <= 2e-5


def test_surface_row_vs_oracle(ctx):
    x, ref_rows = make_rows(5, 4096)
    W = 112
    res = ctx.psd(x, epilogue=True, W=W, want_stats=True, want_db=False)
    mag, _ = ctx.surface_row(res["cols"], res["stats"])
    for f, r in enumerate(ref_rows):
        want, _ = O.surface_row(r, W)
        assert np.max(np.abs(mag[f] - want)) <= 1 and np.mean(mag[f] != want) <= 0.02


def test_display_quantised_planes_vs_oracle(ctx):
    """8f-3: glyph / colour planes (what the reference computes per cell in Python loops)."""
    x, ref_rows = make_rows(32, 4096)
    W, H = 112, 36
    res = ctx.psd(x, epilogue=True, W=W, want_stats=True, want_db=False)
    norm, _ = ctx.display_render(res["cols"], res["stats"], rows_max=30)
    level, colour = ctx.display_quantise(norm[-1], "waterfall")
    hist = []
    for r in ref_rows:
        want_norm, _, want_colour, want_level = O.waterfall_accumulate(hist, r, W)
    assert np.mean(level != want_level) <= 1e-3 and np.mean(colour != want_colour) <= 1e-3
    gnorm, _ = ctx.display_render(res["cols"], res["stats"], rows_max=30, guard_zero_range=True)
    chars, colour = ctx.display_quantise(gnorm[-1], "gradient")
    hist = []
    for r in ref_rows:
        _, _, want_chars, want_colour = O.gradient_accumulate(hist, r, W)
    assert np.mean(chars != want_chars) <= 1e-3 and np.mean(colour != want_colour) <= 1e-3
    # rows older than the history are marked 255
    early, _ = ctx.display_quantise(norm[2], "waterfall")
    assert np.all(early[3:] == 255) and np.all(early[:3] < 4)
    pn, _ = ctx.display_render(res["cols"][:10], res["stats"][:10], rows_max=10, guard_zero_range=True)
    ys, _ = ctx.display_quantise(pn[-1][::-1], "persistence", H=H)
    hist = []
    for r in ref_rows[:10]:
        want_ys, _, _ = O.persistence_accumulate(hist, r, W, H)
    assert np.mean(ys != want_ys) <= 2e-3 and np.max(np.abs(ys.astype(int) - want_ys)) <= 1
