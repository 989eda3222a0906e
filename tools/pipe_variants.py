"""Host->host pipeline time per GiB for several frame lengths, with and without the demodulator."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspecsdr_b200 import core
ctx = core.Context(0)
NB, N, FS = 4096, 32768, 2.4e6
host = ctx.pinned_empty((NB, N), np.complex64)
host.view(np.float32)[:] = np.random.default_rng(0).standard_normal((NB, 2 * N), dtype=np.float32) * 0.5
for n_fft in (4096, 8192, 16384, 32768):
    for mode in ("WFM", None):
        fpb = N // n_fft
        outs = {"audio": ctx.pinned_empty((NB, 304, 2)), "cols": ctx.pinned_empty((NB * fpb, 200)),
                "stats": ctx.pinned_empty((NB * fpb, 4)), "norm": ctx.pinned_empty((NB, 30, 200)),
                "minmax": ctx.pinned_empty((NB, 2))}
        ctx.pipeline(host, FS, mode, n_fft, 200, 30, out=outs)
        t0 = time.perf_counter()
        for _ in range(2):
            ctx.pipeline(host, FS, mode, n_fft, 200, 30, out=outs)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 2 * 1e3
        print(f"n_fft={n_fft} mode={mode}: {ms:.2f} ms per GiB", flush=True)
