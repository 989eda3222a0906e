"""GPU: INTEGER outputs are bit-exact.

* display planes (glyph / colour / screen-row / magnitude) from the stateful display streams, compared
  DIRECTLY with the cells the unmodified reference drew (tests/golden/display.npz), no tolerance;
* the normalised fp64 values behind them, bit-equal to the oracle's numpy arithmetic;
* scanner bin counts against the reference-executed goldens and the oracle, no tolerance;
* the int16 audio pack against the golden, no tolerance;
* statefulness: N calls of one row / block == one call of N rows / blocks, bitwise, with the 34-row wrap.
"""
import numpy as np
import pytest

import _display_cells as D
from oracle import ref_dsp as O
from pyspecsdr_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def disp(golden):
    g = golden("display")
    return g, D.golden_rows(g)


def _feed_one_by_one(ctx, stream, rows):
    """The reference's call pattern: one draw_* call (= one accumulate) per main-loop iteration."""
    outs = [ctx.display_accumulate(stream, r) for r in rows]
    return {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}


def test_waterfall_planes_equal_reference_cells(ctx, disp):
    g, rows = disp
    W = int(g["W"]) - 8
    ctx.display_open(1, "waterfall", W=W, rows_max=30)
    res = _feed_one_by_one(ctx, 1, rows)
    hist = []
    for s, r in enumerate(rows):
        norm, (lo, hi), colour, level = O.waterfall_accumulate(hist, r, W)
        n = len(hist)
        assert res["n_rows"][s] == n
        np.testing.assert_array_equal(res["norm64"][s, :n], norm)                 # fp64, bit for bit
        assert np.all(np.isnan(res["norm64"][s, n:])) and np.all(res["plane_a"][s, n:] == 255)
        np.testing.assert_array_equal(res["minmax64"][s], [lo, hi])
        np.testing.assert_array_equal(res["plane_a"][s, :n], level)
        np.testing.assert_array_equal(res["plane_b"][s, :n], colour)
        if s in (0, 5, 33):                                                       # what the reference drew
            ch, at = D.waterfall_cells(res["plane_a"][s, :n], res["plane_b"][s, :n])
            np.testing.assert_array_equal(ch, g[f"waterfall_{s}_char"])
            np.testing.assert_array_equal(at, g[f"waterfall_{s}_attr"])
    assert ctx.display_rows(1) == 29
    # one call of 34 rows == 34 calls of one row (the ring is carried on the device)
    ctx.display_open(2, "waterfall", W=W, rows_max=30)
    batch = ctx.display_accumulate(2, np.array(rows))
    for k in res:
        np.testing.assert_array_equal(batch[k], res[k], err_msg=k)
    # ... and so is any split
    ctx.display_open(2, "waterfall", W=W, rows_max=30)
    parts = [ctx.display_accumulate(2, np.array(rows[a:b])) for a, b in ((0, 3), (3, 4), (4, 33), (33, 34))]
    for k in res:
        np.testing.assert_array_equal(np.concatenate([p[k] for p in parts]), res[k], err_msg=k)
    ctx.display_close(1)
    ctx.display_close(2)


def test_gradient_planes_equal_reference_cells(ctx, disp):
    g, rows = disp
    W = int(g["W"]) - 10
    ctx.display_open(3, "gradient", W=W, rows_max=30)
    res = _feed_one_by_one(ctx, 3, rows)
    hist = []
    for s, r in enumerate(rows):
        norm, _, chars, colour = O.gradient_accumulate(hist, r, W)
        n = len(hist)
        np.testing.assert_array_equal(res["norm64"][s, :n], norm)
        np.testing.assert_array_equal(res["plane_a"][s, :n], chars)
        np.testing.assert_array_equal(res["plane_b"][s, :n], colour)
        if s in (0, 33):
            ch, at = D.gradient_cells(res["plane_a"][s, :n], res["plane_b"][s, :n])
            np.testing.assert_array_equal(ch, g[f"gradient_{s}_char"])
            np.testing.assert_array_equal(at, g[f"gradient_{s}_attr"])
    ctx.display_close(3)


def test_persistence_planes_equal_reference_stars(ctx, disp):
    g, rows = disp
    H, W = int(g["H"]) - 4, int(g["W"]) - 8
    ctx.display_open(4, "persistence", W=W, rows_max=10, H=H)
    res = _feed_one_by_one(ctx, 4, rows[:14])
    hist = []
    for s, r in enumerate(rows[:14]):
        ys, colours, (lo, hi) = O.persistence_accumulate(hist, r, W, H)
        n = len(hist)
        got_y = res["plane_a"][s, :n][::-1].astype(np.int64)          # oldest trace first, like PERSISTENCE_HISTORY
        got_c = res["plane_b"][s, :n, 0][::-1].astype(np.int64)
        np.testing.assert_array_equal(got_y, ys)
        np.testing.assert_array_equal(got_c, colours)
        np.testing.assert_array_equal(res["minmax64"][s], [lo, hi])
        if s in (0, 13):
            np.testing.assert_array_equal(D.persistence_stars(got_y, got_c, H), g[f"persistence_{s}_stars"])
    ctx.display_close(4)


def test_surface_plane_equals_reference_cells(ctx, disp):
    g, rows = disp
    H, Wt = int(g["H"]), int(g["W"])
    ctx.display_open(5, "surface", W=Wt - 8, rows_max=1)
    res = ctx.display_accumulate(5, np.array(rows[:6]))
    for s in range(6):
        mag, (lo, hi) = O.surface_row(rows[s], Wt - 8)
        np.testing.assert_array_equal(res["plane_a"][s, 0], mag)
        np.testing.assert_array_equal(res["minmax64"][s], [lo, hi])
    want = {(int(a), int(b)) for a, b, _ in g["surface_hash_cells"]}
    assert D.surface_cells(res["plane_a"][3, 0], H, Wt) == want
    ctx.display_close(5)


def test_spectrum_view_equals_reference_cells(ctx, disp):
    g, rows = disp
    H, Wt = int(g["H"]), int(g["W"])
    dh, dw = H - 4, Wt - 7
    cols, rng = ctx.spectrum_normalise(np.array(rows[:6]), dw)
    assert cols.dtype == np.float64
    mism = 0
    for s in range(6):
        want_cols, (dmin, dmax) = O.spectrum_normalise(rows[s], dw)
        np.testing.assert_array_equal(rng[s], [dmin, dmax])                      # percentile, range: exact
        assert np.max(np.abs(cols[s] - want_cols)) <= 4 * np.finfo(np.float64).eps   # pow(): <= 2 ulp apart
        mism += int(np.sum(np.minimum((cols[s] * dh).astype(int), dh) != np.minimum((want_cols * dh).astype(int), dh)))
    assert mism == 0                                                             # measured: 0 bar heights differ
    want = {(int(y), int(x)): (int(c), int(a)) for y, x, c, a in g["spectrum_cells"]}
    got = D.spectrum_cells(cols[3], dh)
    assert len(got) == dh * dw and all(want[k] == v for k, v in got.items())


def test_scanner_count_is_exact(ctx, golden):
    g = golden("scanner")
    fr = synth.scanner_frames(24, 2048, seed=3)
    peak, count = ctx.scan(fr, rel_db=20.0)
    np.testing.assert_array_equal(count, g["count"])
    np.testing.assert_array_equal(count * (2.4e6 / 2048), g["bandwidth"])        # :2552, fp64 on the host
    fr8 = synth.scanner_frames(6, 8192, seed=4)
    peak8, count8 = ctx.scan(fr8)
    np.testing.assert_array_equal(count8, g["count8k"])
    assert np.max(np.abs(peak - g["peak"])) <= 1e-4 and np.max(np.abs(peak8 - g["peak8k"])) <= 1e-4
    # absolute-threshold mask (scan_frequencies, pyspecsdr.py:1055-1057)
    _, cabs = ctx.scan(fr, threshold=-40.0)
    want = [O.scan_step(f, 2.4e6, threshold=-40.0)[1] for f in fr]
    np.testing.assert_array_equal(cabs, want)
    # C4-shaped sweep: 8192-point steps, every count equal to the oracle's
    sweep = synth.scanner_frames(96, 8192, seed=8)
    _, cnt = ctx.scan(sweep)
    np.testing.assert_array_equal(cnt, [O.scan_step(f, 2.4e6)[1] for f in sweep])


def test_int16_pack_is_exact(ctx, golden):
    g = golden("int16")
    assert g["audio"].dtype == np.float64
    pcm = ctx.to_int16(g["audio"])
    assert pcm.dtype == np.int16
    np.testing.assert_array_equal(pcm, g["pcm"])
    # truncation toward zero, both signs, and the float32 entry point on float32-representable input
    a = np.array([0.99999, -0.99999, 1.0, -1.0, 0.5 / 32767, -0.5 / 32767, 1.5 / 32767, -1.5 / 32767, 0.0, -0.0])
    np.testing.assert_array_equal(ctx.to_int16(a), np.int16(a * 32767))
    a32 = g["audio"].astype(np.float32)
    np.testing.assert_array_equal(ctx.to_int16(a32), np.int16(a32.astype(np.float64) * 32767))


def test_pipeline_display_ring_is_carried_across_calls(ctx):
    """pss_pipeline_c64 with a display stream: N calls of one block == one call of N blocks, bitwise
    (norm, minmax, planes), across the ring wrap and the pipeline's copy chunks; and a fresh stream equals
    the stateless history of one call."""
    n_block, n_fft, W, R = 4096, 1024, 64, 30
    fs = 1.024e6
    base = np.stack([synth.make("wbfm" if s % 3 else "tone40", n_block, seed=s, fs=fs) if s % 3 else
                     synth.make("tone40", n_block, seed=s) for s in range(12)])
    blocks = np.ascontiguousarray(np.tile(base, (25, 1))[:290])            # 290 blocks: 2 copy chunks
    ctx.display_open(7, "waterfall", W=W, rows_max=R)
    one = ctx.pipeline(blocks, fs, "NFM", n_fft=n_fft, W=W, rows_max=R, display_stream=7, want_planes=True)
    stateless = ctx.pipeline(blocks, fs, "NFM", n_fft=n_fft, W=W, rows_max=R)
    np.testing.assert_array_equal(one["norm"], stateless["norm"])
    np.testing.assert_array_equal(one["minmax"], stateless["minmax"])
    ctx.display_open(7, "waterfall", W=W, rows_max=R)                        # reset
    acc = {k: [] for k in ("norm", "minmax", "plane_a", "plane_b", "audio")}
    for a, b in [(i, i + 1) for i in range(40)] + [(40, 47), (47, 290)]:
        r = ctx.pipeline(blocks[a:b], fs, "NFM", n_fft=n_fft, W=W, rows_max=R, display_stream=7, want_planes=True)
        for k in acc:
            acc[k].append(r[k].copy())
    for k in acc:
        np.testing.assert_array_equal(np.concatenate(acc[k]), one[k], err_msg=k)
    # planes are the quantisation of the fp64 normalised value; norm is its float32 rounding
    v = one["norm"][-1].astype(np.float64)
    assert np.mean(one["plane_b"][-1] == (v * 5).astype(np.int64)) > 0.999
    ctx.display_close(7)
    from pyspecsdr_b200.core import PssError
    with pytest.raises(PssError):
        ctx.pipeline(blocks[:1], fs, "NFM", n_fft=n_fft, W=W, rows_max=R, display_stream=7)     # closed stream
