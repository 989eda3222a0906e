"""The reference's real main-loop shape (pyspecsdr.py:2236-2283): ONE FFT over the whole 32768-sample read +
smoothing / median clamp + WFM demod of the same read, through the host-pointer pipeline (pinned buffers,
copies inside) and device-resident.  Not the BASELINE bench line (that is C2: 4096-point frames); recorded
for DESIGN.md."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspecsdr_b200 import core
ctx = core.Context(0)
NB, N, FS = 4096, 32768, 2.4e6
host = ctx.pinned_empty((NB, N), np.complex64)
host.view(np.float32)[:] = np.random.default_rng(0).standard_normal((NB, 2 * N), dtype=np.float32) * 0.5
for n_fft in (32768, 4096):
    fpb = N // n_fft
    outs = {"audio": ctx.pinned_empty((NB, 304, 2)), "cols": ctx.pinned_empty((NB * fpb, 200)),
            "stats": ctx.pinned_empty((NB * fpb, 4)), "norm": ctx.pinned_empty((NB, 30, 200)),
            "minmax": ctx.pinned_empty((NB, 2))}
    for _ in range(2):                       # first calls allocate scratch and load the kernels
        ctx.pipeline(host, FS, "WFM", n_fft, 200, 30, out=outs)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.pipeline(host, FS, "WFM", n_fft, 200, 30, out=outs)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print(f"pipeline host->host n_fft={n_fft}: {ms:.2f} ms per GiB = {NB * N / ms / 1e3:.0f} MS/s", flush=True)
# device-resident
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
iq = torch.from_numpy(host.view(np.float32).reshape(NB, N, 2)).cuda()
plan = ctx.demod_plan("WFM", FS, N)
audio = torch.empty(NB, plan.out_len, plan.channels, device="cuda")
for n_fft in (32768, 4096):
    F = NB * (N // n_fft)
    db = torch.empty(F, n_fft - 4, device="cuda"); cols = torch.empty(F, 200, device="cuda")
    stats = torch.empty(F, 4, device="cuda"); mom = torch.empty(F, 4, device="cuda", dtype=torch.float64)
    def step():
        ctx.psd_dev(iq, n_fft, F, db=db, window="hamming", epilogue=True, cols=cols, W=200, stats=stats, moments=mom)
        ctx.demod_dev(plan, iq, NB, audio, moments=mom, frames_per_block=N // n_fft)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"device-resident psd+epilogue+WFM n_fft={n_fft}: {ms:.3f} ms per GiB = {NB * N / ms / 1e3:.0f} MS/s", flush=True)
