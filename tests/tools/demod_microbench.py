"""Scratch micro-benchmark of the demodulation kernels (device-resident input, CUDA events)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from pyspecsdr_b200 import core
from oracle import ref_dsp as O

ctx = core.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
N = 32768
F = 4096
modes = sys.argv[1:] or ["NFM", "WFM"]
for mode in modes:
    fs = 2.4e6 if mode in ("NFM", "WFM") else 1e6
    plan = ctx.demod_plan(mode, fs, N)
    # FM-like synthetic signal generated on the device
    t = torch.arange(N, device="cuda", dtype=torch.float64) / fs
    ph = 2 * np.pi * 75e3 * torch.cumsum(torch.sin(2 * np.pi * 1e3 * t), 0) / fs
    base = torch.stack([torch.cos(ph), torch.sin(ph)], -1).to(torch.float32)
    iq = base[None].repeat(F, 1, 1) + 0.01 * torch.randn(F, N, 2, device="cuda")
    audio = torch.empty(F, plan.out_len, plan.channels, device="cuda", dtype=torch.float32)
    run = lambda: ctx.demod_dev(plan, iq, F, audio)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_alg = F * N * 8 + audio.numel() * 4
    # parity on a few frames
    xs = iq[:2].cpu().numpy()
    xs = (xs[..., 0] + 1j * xs[..., 1]).astype(np.complex64)
    got = audio[:2].cpu().numpy()
    errs = []
    for f in range(2):
        ref = O.demod(xs[f], fs, mode)
        mono = ref[:, 0] if ref.ndim == 2 else ref
        errs.append(float(np.sqrt(np.mean((got[f][:, 0] - mono) ** 2))))
    print(f"{mode}: {ms:.3f} ms  {F*N/ms/1e3:.1f} MS/s  {bytes_alg/ms/1e6:.0f} GB/s algorithmic "
          f"frac={bytes_alg/ms/1e6/peak:.3f}  rms_err={max(errs):.2e}", flush=True)
