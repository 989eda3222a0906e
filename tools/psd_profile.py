"""A few launches of the fused PSD kernel for ncu (N from argv, default 4096)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspecsdr_b200 import core
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = core.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
total = 1 << 27
iq = torch.randn(total, 2, device="cuda", dtype=torch.float32)
F = total // N
db = torch.empty(F * N, device="cuda", dtype=torch.float32)
cols = torch.empty(F * 200, device="cuda", dtype=torch.float32)
stats = torch.empty(F * 4, device="cuda", dtype=torch.float32)
for _ in range(2):
    ctx.psd_dev(iq, N, F, db=db)
    ctx.psd_dev(iq, N, F, db=db, epilogue=True, cols=cols, W=200, stats=stats)
torch.cuda.synchronize()
