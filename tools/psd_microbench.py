"""Scratch micro-benchmark of the fused PSD kernel (device-resident input, CUDA events)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspecsdr_b200 import core

ctx = core.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
total = 1 << 27          # 128 Mi samples = 1 GiB of complex64
iq = torch.randn(total, 2, device="cuda", dtype=torch.float32)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
for N in (1024, 2048, 4096, 8192):
    F = total // N
    for epi in (False, True, "fp32"):
        fp32 = epi == "fp32"
        epi = epi is True
        n_out = N - 4 if epi else N
        db = torch.empty(F * n_out, device="cuda", dtype=torch.float32)
        cols = torch.empty(F * 200, device="cuda", dtype=torch.float32) if epi else None
        stats = torch.empty(F * 4, device="cuda", dtype=torch.float32) if epi else None
        run = lambda: ctx.psd_dev(iq, N, F, db=db, epilogue=epi, cols=cols, W=200 if epi else 0, stats=stats, fp32=fp32)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = total * 12 / ms / 1e6
        print(f"N={N} epilogue={epi} fp32={fp32}: {ms:.3f} ms  {total/ms/1e3:.1f} MS/s  {gbs:.0f} GB/s algorithmic  "
              f"frac={gbs/peak:.3f}", flush=True)
