"""Does a concurrent 64 MB host->device copy slow the per-chunk kernels of the host pipeline (or vice versa)?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspecsdr_b200 import core
ctx = core.Context(0)
comp = torch.cuda.Stream(); cp = torch.cuda.Stream()
ctx.set_stream(comp.cuda_stream)
N, nb = 32768, 256
host = torch.empty(nb, N, 2, dtype=torch.float32).pin_memory()
host.normal_()
dst = torch.empty(nb, N, 2, device="cuda")
iq = (torch.randn(nb, N, 2, device="cuda") * 0.5)
for n_fft in (32768, 4096):
    F = nb * (N // n_fft)
    db = torch.empty(F, n_fft - 4, device="cuda"); cols = torch.empty(F, 200, device="cuda"); stats = torch.empty(F, 4, device="cuda")
    def psd(): ctx.psd_dev(iq, n_fft, F, db=db, window="hamming", epilogue=True, cols=cols, W=200, stats=stats)
    for _ in range(3): psd()
    torch.cuda.synchronize()
    for with_copy in (False, True):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if with_copy:
            with torch.cuda.stream(cp):
                c0.record(cp)
                for _ in range(16): dst.copy_(host, non_blocking=True)
                c1.record(cp)
        e0.record(comp)
        for _ in range(16): psd()
        e1.record(comp)
        torch.cuda.synchronize()
        msg = f"n_fft={n_fft} copy={with_copy}: psd {e0.elapsed_time(e1)/16:.3f} ms per launch"
        if with_copy: msg += f", copy {c0.elapsed_time(c1)/16:.3f} ms per 64 MB"
        print(msg, flush=True)
