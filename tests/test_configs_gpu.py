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
        assert abs(float(peak[k]) - pk) <= 1e-4 and abs(int(count[k]) - cnt) <= 1


def test_c5_persistence_surface_16384(ctx):
    """C5: 16384-pt frames, 10-row persistence history + surface row, W = 200."""
    n, W, H = 16384, 200, 36
    x = np.stack([synth.make("wbfm" if s % 2 else "tone40", n, seed=200 + s) for s in range(13)])
    res = ctx.psd(x, epilogue=True, W=W, want_stats=True, want_db=False)
    rows = [O.psd_epilogue(O.psd_db(r)) for r in x]
    norm, mm = ctx.display_render(res["cols"], res["stats"], rows_max=10, guard_zero_range=True)
    hist = []
    for s, r in enumerate(rows):
        ys, colours, (lo, hi) = O.persistence_accumulate(hist, r, W, H)
        got = norm[s, :len(hist)][::-1]
        y = ((1 - got.astype(np.float64)) * (H - 1)).astype(np.int64)
        assert np.max(np.abs(y - ys)) <= 1 and np.mean(y != ys) <= 2e-3
        assert abs(mm[s, 0] - lo) <= 1e-4 and abs(mm[s, 1] - hi) <= 1e-4
    # surface row = the frame normalised by its own range (history of one row), magnitude = int(v * 20)
    snorm, _ = ctx.display_render(res["cols"], res["stats"], rows_max=1, guard_zero_range=True)
    for s in (0, 7, 12):
        mag, _ = O.surface_row(rows[s], W)
        got = (snorm[s, 0].astype(np.float64) * 20).astype(np.int64)
        assert np.max(np.abs(got - mag)) <= 1 and np.mean(got != mag) <= 0.02


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
