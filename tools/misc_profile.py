"""A few launches of the kernels that are not part of the bench step (for ncu): large-transform PSD with
the epilogue, SSB FIR, Welch + classify."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspecsdr_b200 import core
ctx = core.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
N, F = 32768, 2048
iq = torch.randn(F, N, 2, device="cuda", dtype=torch.float32)
db = torch.empty(F * (N - 4), device="cuda")
cols = torch.empty(F * 200, device="cuda")
stats = torch.empty(F * 4, device="cuda")
plan = ctx.demod_plan("USB", 1e6, N)
audio = torch.empty(F, plan.out_len, plan.channels, device="cuda")
feat = torch.empty(F, 4, device="cuda", dtype=torch.float64)
lab = torch.empty(F, device="cuda", dtype=torch.int32)
for _ in range(2):
    ctx.psd_dev(iq, N, F, db=db, epilogue=True, cols=cols, W=200, stats=stats)
    ctx.demod_dev(plan, iq, F, audio)
    ctx.classify_dev(iq, N, F, 2.4e6, feat, lab)
torch.cuda.synchronize()
