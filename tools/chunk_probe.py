import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspecsdr_b200 import core
ctx = core.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
N, FS = 32768, 2.4e6
plan = ctx.demod_plan("WFM", FS, N)
for nb in (256, 4096):
    iq = torch.randn(nb, N, 2, device="cuda") * 0.5
    audio = torch.empty(nb, plan.out_len, plan.channels, device="cuda")
    for n_fft in (32768, 4096):
        F = nb * (N // n_fft)
        db = torch.empty(F, n_fft - 4, device="cuda"); cols = torch.empty(F, 200, device="cuda")
        stats = torch.empty(F, 4, device="cuda"); mom = torch.empty(F, 4, device="cuda", dtype=torch.float64)
        def psd(): ctx.psd_dev(iq, n_fft, F, db=db, window="hamming", epilogue=True, cols=cols, W=200, stats=stats, moments=mom)
        def dem(): ctx.demod_dev(plan, iq, nb, audio, moments=mom, frames_per_block=N // n_fft)
        for name, fn in (("psd", psd), ("demod", dem)):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            for _ in range(16): fn()
            e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
            print(f"nb={nb} n_fft={n_fft} {name}: {e0.elapsed_time(e1)/16:.3f} ms per launch (host enqueue {1e3*(t1-t0)/16:.3f} ms)", flush=True)
