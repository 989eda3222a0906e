"""A few launches of the demod kernel for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspecsdr_b200 import core
mode = sys.argv[1] if len(sys.argv) > 1 else "WFM"
ctx = core.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
N, F = 32768, 2048
fs = 2.4e6 if mode in ("NFM", "WFM") else 1e6
plan = ctx.demod_plan(mode, fs, N)
iq = torch.randn(F, N, 2, device="cuda", dtype=torch.float32)
audio = torch.empty(F, plan.out_len, plan.channels, device="cuda", dtype=torch.float32)
for _ in range(3):
    ctx.demod_dev(plan, iq, F, audio)
torch.cuda.synchronize()
