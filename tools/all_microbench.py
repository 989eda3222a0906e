"""Scratch micro-benchmark: every PSD size (raw / smoothing epilogue) and every demod mode,
device-resident 1 GiB inputs, CUDA events.  Prints one line per case and writes a JSON table."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyspecsdr_b200 import core

ctx = core.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
total = 1 << 27
iq = torch.randn(total, 2, device="cuda", dtype=torch.float32)
rows = []

def timeit(run, reps=10):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

sizes = [int(a) for a in os.environ.get("PSD_SIZES", "1024,2048,4096,8192,16384,32768,65536,131072").split(",") if a]
for N in sizes:
    F = total // N
    for epi in (False, True):
        n_out = N - 4 if epi else N
        db = torch.empty(F * n_out, device="cuda", dtype=torch.float32)
        cols = torch.empty(F * 200, device="cuda", dtype=torch.float32) if epi else None
        stats = torch.empty(F * 4, device="cuda", dtype=torch.float32) if epi else None
        ms = timeit(lambda: ctx.psd_dev(iq, N, F, db=db, epilogue=epi, cols=cols, W=200 if epi else 0, stats=stats))
        gbs = total * 12 / ms / 1e6
        rows.append({"kernel": "psd", "N": N, "epilogue": epi, "ms": ms, "GSps": total / ms / 1e6, "frac": gbs / peak})
        print(rows[-1], flush=True)
    # scanner
    pk = torch.empty(F, device="cuda", dtype=torch.float32)
    cnt = torch.empty(F, device="cuda", dtype=torch.int32)
    if N <= 8192 and N >= 512 and hasattr(ctx, "scan_dev"):
        ms = timeit(lambda: ctx.scan_dev(iq, N, F, pk, cnt, rel_db=20.0))
        rows.append({"kernel": "scan", "N": N, "ms": ms, "GSps": total / ms / 1e6, "frac": total * 8 / ms / 1e6 / peak})
        print(rows[-1], flush=True)

N = 32768
F = 4096
for mode in os.environ.get("DEMOD_MODES", "NFM,WFM,AM,USB,RAW").split(","):
    if not mode:
        continue
    fs = 2.4e6 if mode in ("NFM", "WFM") else 1e6
    plan = ctx.demod_plan(mode, fs, N)
    audio = torch.empty(F, plan.out_len, plan.channels, device="cuda", dtype=torch.float32)
    x = iq.view(F, N, 2)
    ms = timeit(lambda: ctx.demod_dev(plan, x, F, audio), reps=5)
    b = F * N * 8 + audio.numel() * 4
    rows.append({"kernel": "demod", "mode": mode, "ms": ms, "GSps": F * N / ms / 1e6, "frac": b / ms / 1e6 / peak})
    print(rows[-1], flush=True)
if os.environ.get("CLASSIFY", "1") == "1":
    feat = torch.empty(F, 4, device="cuda", dtype=torch.float64)
    lab = torch.empty(F, device="cuda", dtype=torch.int32)
    x = iq.view(F, N, 2)
    ms = timeit(lambda: ctx.classify_dev(x, N, F, 2.4e6, feat, lab), reps=5)
    rows.append({"kernel": "classify", "ms": ms, "GSps": F * N / ms / 1e6, "frac": F * N * 8 / ms / 1e6 / peak})
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/all_microbench.json", "w"), indent=1)
