"""GPU parity: display accumulate fed by the PSD kernel's float32 rows (IQ in -> display values out) vs the
oracle, which is itself pinned through what the reference's draw_* functions drew.

Floating-point outputs: within the stated tolerance.  Integer planes on this path are derived from spectra
that carry the 1e-4 dB tolerance, so they must EQUAL the oracle's except in cells where the oracle's own
value lies on a quantisation boundary (`assert_equal_or_on_boundary`; the count of such cells is bounded).
The planes themselves are bit-exact given the same dB rows: tests/test_exact_gpu.py checks that against the
cells the reference drew."""
import numpy as np
import pytest

import _display_cells as D
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
    for s, r in enumerate(ref_rows):
        want, (lo, hi), colour, level = O.waterfall_accumulate(hist, r, W)
        got = norm[s, :len(hist)]
        assert abs(mm[s, 0] - lo) <= 1e-4 and abs(mm[s, 1] - hi) <= 1e-4
        assert np.max(np.abs(got - want)) <= 1e-5
        assert np.all(np.isnan(norm[s, len(hist):]))


@pytest.mark.parametrize("W,rows_max", [(200, 30), (113, 30), (66, 7), (4, 2), (1028, 31)])
def test_render_widths_vs_oracle(ctx, W, rows_max):
    """Both forms of the values-only render (16-byte path for W % 4 == 0, scalar otherwise) against the oracle,
    with NaN padding for the rows the history does not hold yet."""
    x, ref_rows = make_rows(rows_max + 3, 2048)
    res = ctx.psd(x, epilogue=True, W=W, want_stats=True, want_db=False)
    norm, mm = ctx.display_render(res["cols"], res["stats"], rows_max=rows_max)
    hist = []
    for s, r in enumerate(ref_rows):
        want, (lo, hi), _, _ = O.waterfall_accumulate(hist, r, W, max_rows=rows_max)
        assert np.max(np.abs(norm[s, :len(hist)] - want)) <= 1e-5
        assert np.all(np.isnan(norm[s, len(hist):]))
        assert abs(mm[s, 0] - lo) <= 1e-4 and abs(mm[s, 1] - hi) <= 1e-4


def test_persistence_history_vs_oracle(ctx):
    x, ref_rows = make_rows(14)
    W, H = 112, 36
    res = ctx.psd(x, epilogue=True, W=W, want_stats=True, want_db=False)
    norm, mm = ctx.display_render(res["cols"], res["stats"], rows_max=10, guard_zero_range=True)
    hist = []
    for s, r in enumerate(ref_rows):
        ys, colours, (lo, hi) = O.persistence_accumulate(hist, r, W, H)
        got = norm[s, :len(hist)][::-1]                      # oldest first, like PERSISTENCE_HISTORY
        want = np.stack([(O.resample_cols(line, W) - lo) / (hi - lo) for line in hist])
        assert np.max(np.abs(got - want)) <= 1e-5


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


TOL_NORM = 2e-4 / 40.0      # 2 x the dB tolerance over a >= 40 dB display range, in normalised units


def test_planes_from_iq_equal_oracle_except_on_boundaries(ctx):
    """8f-3 on the product path: IQ -> PSD kernel (float32 rows on the device) -> carried display streams
    -> uint8 planes (what the reference computes per cell in Python loops)."""
    import torch
    x, ref_rows = make_rows(32, 4096)
    W, H = 112, 36
    dev = torch.device("cuda", 0)
    xd = torch.from_numpy(x.view(np.float32).reshape(32, 4096, 2)).to(dev)
    cols = torch.empty(32, W, device=dev)
    stats = torch.empty(32, 4, device=dev)
    planes = {k: (torch.empty(32, r, W, device=dev, dtype=torch.uint8), torch.empty(32, r, W, device=dev, dtype=torch.uint8))
              for k, r in (("waterfall", 30), ("gradient", 30), ("persistence", 10), ("surface", 1))}
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        ctx.psd_dev(xd, 4096, 32, epilogue=True, cols=cols, W=W, stats=stats)
        for i, (k, (pa, pb)) in enumerate(planes.items()):
            ctx.display_open(30 + i, k, W=W, rows_max=pa.shape[1], H=H if k == "persistence" else 0)
            ctx.display_accumulate_dev(30 + i, cols, stats, 32, plane_a=pa, plane_b=pb)
            ctx.display_close(30 + i)
        torch.cuda.synchronize()
    finally:
        ctx.set_stream(None)
    P = {k: (a.cpu().numpy(), b.cpu().numpy()) for k, (a, b) in planes.items()}
    hw, hp, nb = [], [], 0
    for s, r in enumerate(ref_rows):
        want_norm, (lo, hi), _, want_level = O.waterfall_accumulate(hw, r, W)
        L = len(hw)
        assert hi - lo >= 40.0
        nb += D.assert_equal_or_on_boundary(P["waterfall"][1][s, :L], want_norm, 5, TOL_NORM, "waterfall colour")
        bad = P["waterfall"][0][s, :L] != want_level                      # '.', '-', '=', '#': strict > 0.25, 0.5, 0.75
        if bad.any():
            assert np.all(np.min(np.abs(want_norm[bad][:, None] - np.array([0.25, 0.5, 0.75])), axis=1) <= TOL_NORM)
            nb += int(bad.sum())
        nb += D.assert_equal_or_on_boundary(P["gradient"][0][s, :L], want_norm, 8, TOL_NORM, "gradient glyph")
        nb += D.assert_equal_or_on_boundary(P["gradient"][1][s, :L], want_norm, 5, TOL_NORM, "gradient colour")
        assert np.all(P["waterfall"][0][s, L:] == 255)
        ys, colours, (plo, phi) = O.persistence_accumulate(hp, r, W, H)
        Lp = len(hp)
        pn = np.stack([(O.resample_cols(line, W) - plo) / (phi - plo) for line in hp])
        nb += D.assert_equal_or_on_boundary(P["persistence"][0][s, :Lp][::-1], 1 - pn, H - 1, TOL_NORM, "persistence y")
        np.testing.assert_array_equal(P["persistence"][1][s, :Lp, 0][::-1], colours)
        fin = r[np.isfinite(r)]
        surf = O.resample_cols((r - fin.min()) / (fin.max() - fin.min()), W)
        nb += D.assert_equal_or_on_boundary(P["surface"][0][s, 0], surf, 20, TOL_NORM, "surface magnitude")
    assert nb <= 2e-4 * 32 * 72 * W          # a handful of boundary cells out of ~258 000


def test_quantise_call_matches_stream_planes(ctx):
    """pss_display_quantise (planes from already-normalised float32 values) uses the same rules."""
    v = np.linspace(-0.1, 1.1, 1201).astype(np.float32)
    v[7] = np.nan
    d = np.nan_to_num(v.astype(np.float64))
    a, b = ctx.display_quantise(v, "waterfall")
    ok = np.isfinite(v)
    np.testing.assert_array_equal(a[ok], ((d > 0.25).astype(int) + (d > 0.5) + (d > 0.75))[ok])
    np.testing.assert_array_equal(b[ok], np.clip((d * 5).astype(np.int64), 0, 254)[ok])
    assert a[7] == 255 and b[7] == 255
    a, _ = ctx.display_quantise(v, "gradient")
    np.testing.assert_array_equal(a[ok], np.clip((d * 8).astype(np.int64), 0, 254)[ok])
    a, _ = ctx.display_quantise(v, "surface")
    np.testing.assert_array_equal(a[ok], np.clip((d * 20).astype(np.int64), 0, 254)[ok])
    a, _ = ctx.display_quantise(v, "persistence", H=36)
    y = ((1 - d) * 35).astype(np.int64)
    want = np.where((y >= 0) & (y < 36), y, 255)
    np.testing.assert_array_equal(a[ok], want[ok])
