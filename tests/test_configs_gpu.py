"""GPU: the BASELINE.json configurations that are parity cases rather than bench lines (C3-C5), at
reduced batch counts but full per-unit sizes."""
import numpy as np
import pytest

from oracle import ref_dsp as O
from pyspecsdr_b200 import shard, synth

pytestmark = pytest.mark.gpu


def test_c3_am_ssb_1msps(ctx):
    """C3: AM + USB/LSB demod chain at 1 MS/s on 32768-sample blocks."""
    fs, n = 1e6, 32768
    am = np.stack([synth.am_tone(n, seed=s, fs=fs) for s in range(4)])
    ssb = np.stack([synth.ssb_two_tone(n, seed=s, fs=fs) for s in range(4)])
    for mode, x in (("AM", am), ("USB", ssb), ("LSB", ssb)):
        got = ctx.demod(x, fs, mode)
        for f in range(len(x)):
            ref = O.demod(x[f], fs, mode)[:, 0]
            assert np.sqrt(np.mean((got[f, :, 0] - ref) ** 2)) <= 1e-5, mode


def test_c4_scanner_sweep_8192_sharded_stitch(ctx):
    """C4: 8192-pt un-windowed PSDs per scan step, (peak, count) records and the stitched dB rows.
    Sharding is exercised through shard.sweep on one rank (world 1) and by processing the sweep as
    8 frame ranges: the stitched result must be bitwise equal to the single pass."""
    import torch
    n_steps, N = 40, 8192
    frames = synth.scanner_frames(n_steps, N, seed=8)
    dev = torch.device("cuda", 0)
    fr_dev = torch.from_numpy(frames.view(np.float32).reshape(n_steps, N, 2)).to(dev)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        peak, count, rows = shard.sweep(ctx, fr_dev, N, n_steps, 0, 1, want_rows=True)
        torch.cuda.synchronize()
        parts = []
        for r in range(8):
            lo, hi = shard.frame_range(n_steps, r, 8)
            p = torch.empty(hi - lo, device=dev)
            c = torch.empty(hi - lo, device=dev, dtype=torch.int32)
            rw = torch.empty(hi - lo, N, device=dev)
            if hi > lo:
                ctx.scan_dev(fr_dev[lo:hi].contiguous(), N, hi - lo, p, c, rows=rw)
            parts.append((p, c, rw))
        torch.cuda.synchronize()
    finally:
        ctx.set_stream(None)
    assert torch.equal(torch.cat([p[0] for p in parts]), peak)
    assert torch.equal(torch.cat([p[1] for p in parts]), count)
    assert torch.equal(torch.cat([p[2] for p in parts]), rows)
    want = O.psd_db(frames, window="none")
    assert np.max(np.abs(rows.cpu().numpy() - want)) <= 1e-4
    for k in range(n_steps):
        pk, cnt, _ = O.scan_step(frames[k], 2.4e6)
        assert abs(float(peak[k]) - pk) <= 1e-4 and int(count[k]) == cnt


def test_c5_persistence_surface_16384(ctx):
    """C5: 16384-pt frames, 10-row persistence history + surface row, W = 200, on a carried display
    stream fed straight from the PSD kernel's device rows (one stream per IQ stream and view).
    The planes come from float32 spectra (1e-4 dB tolerance), so a cell may differ from the oracle's only
    where the oracle's own value sits on a quantisation boundary; everything else must be equal."""
    import torch
    import _display_cells as D
    n, W, H, R = 16384, 200, 36, 10
    x = np.stack([synth.make("wbfm" if s % 2 else "tone40", n, seed=200 + s) for s in range(13)])
    rows = [O.psd_epilogue(O.psd_db(r)) for r in x]
    dev = torch.device("cuda", 0)
    xd = torch.from_numpy(x.view(np.float32).reshape(13, n, 2)).to(dev)
    cols = torch.empty(13, W, device=dev)
    stats = torch.empty(13, 4, device=dev)
    ys = torch.empty(13, R, W, device=dev, dtype=torch.uint8)
    cp = torch.empty(13, R, W, device=dev, dtype=torch.uint8)
    mm = torch.empty(13, 2, device=dev)
    mag = torch.empty(13, 1, W, device=dev, dtype=torch.uint8)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        ctx.display_open(20, "persistence", W=W, rows_max=R, H=H)
        ctx.display_open(21, "surface", W=W, rows_max=1)
        for a, b in ((0, 1), (1, 2), (2, 9), (9, 13)):         # the ring is carried from call to call
            ctx.psd_dev(xd[a:b], n, b - a, epilogue=True, cols=cols[a:b], W=W, stats=stats[a:b])
            ctx.display_accumulate_dev(20, cols[a:b], stats[a:b], b - a, plane_a=ys[a:b], plane_b=cp[a:b], minmax=mm[a:b])
            ctx.display_accumulate_dev(21, cols[a:b], stats[a:b], b - a, plane_a=mag[a:b])
        torch.cuda.synchronize()
    finally:
        ctx.set_stream(None)
        ctx.display_close(20)
        ctx.display_close(21)
    ys, cp, mm, mag = ys.cpu().numpy(), cp.cpu().numpy(), mm.cpu().numpy(), mag.cpu().numpy()
    hist, boundary = [], 0
    tol = 2e-4 / 40.0                       # 2 x the dB tolerance over a >= 40 dB stack range, in normalised units
    for s, r in enumerate(rows):
        ref_y, colours, (lo, hi) = O.persistence_accumulate(hist, r, W, H)
        L = len(hist)
        assert hi - lo >= 40.0 and abs(mm[s, 0] - lo) <= 1e-4 and abs(mm[s, 1] - hi) <= 1e-4
        ref_norm = np.stack([(O.resample_cols(line, W) - lo) / (hi - lo) for line in hist])
        got = ys[s, :L][::-1]
        boundary += D.assert_equal_or_on_boundary(got, 1 - ref_norm, H - 1, tol, f"persistence y, frame {s}")
        np.testing.assert_array_equal(cp[s, :L, 0][::-1], colours)
        assert np.all(ys[s, L:] == 255)
        fin = r[np.isfinite(r)]
        rl, rh = fin.min(), fin.max()
        ref_surf = O.resample_cols((r - rl) / (rh - rl), W)
        boundary += D.assert_equal_or_on_boundary(mag[s, 0], ref_surf, 20, tol, f"surface, frame {s}")
    assert boundary <= 13 * (R + 1) * W * 1e-3


def test_capture_file_processing_sharded(ctx, tmp_path):
    """8f-1: a .npy capture (the reference's record_signal format) processed block by block; two
    'ranks' processing disjoint shares reproduce the single-rank result bitwise (audio, spectra) and
    match the oracle."""
    from pyspecsdr_b200 import capture
    block, n_fft, fs = 8192, 1024, 1.024e6
    x = np.concatenate([synth.wbfm(block, seed=s, fs=fs) for s in range(9)] + [synth.noise(100, 1)])
    path = str(tmp_path / "capture.npy")
    np.save(path, x)
    full = capture.process_capture(ctx, path, fs, "NFM", block=block, n_fft=n_fft, W=64, chunk_blocks=4)
    assert full["audio"].shape[0] == 9
    parts = [capture.process_capture(ctx, path, fs, "NFM", block=block, n_fft=n_fft, W=64, rank=r, world=2,
                                     chunk_blocks=2) for r in range(2)]
    np.testing.assert_array_equal(np.concatenate([p["audio"] for p in parts]), full["audio"])
    np.testing.assert_array_equal(np.concatenate([p["cols"] for p in parts]), full["cols"])
    np.testing.assert_array_equal(np.concatenate([p["stats"] for p in parts]), full["stats"])
    for b in (0, 8):
        ref = O.demod(x[b * block:(b + 1) * block], fs, "NFM")
        assert np.sqrt(np.mean((full["audio"][b] - ref) ** 2)) <= 1e-5
