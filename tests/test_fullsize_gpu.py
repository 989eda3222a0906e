"""GPU: BASELINE.json's full bench sizes (1 GiB of IQ per launch) through size-independent properties.
The oracle cannot run a GiB in seconds, so the batch is built from a few distinct frames / blocks repeated
across the grid: every copy must be bitwise equal to the first occurrence wherever it lands in the launch
(no cross-frame interference, no dependence on CTA / wave placement), and the first occurrences are checked
against the oracle."""
import numpy as np
import pytest

from oracle import ref_dsp as O
from pyspecsdr_b200 import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x).view(np.float32).reshape(x.shape + (2,))).cuda()


def test_psd_epilogue_full_bench_batch(ctx):
    N, F, K = 4096, 32768, 37                       # 37 distinct frames, co-prime with CTAs per SM / SM count
    kinds = ["wbfm", "tone40", "noise", "halfband", "tone60"]
    base = np.stack([synth.make(kinds[i % 5], N, seed=100 + i) for i in range(K)])
    iq = _dev(base).repeat((F + K - 1) // K, 1, 1)[:F].contiguous()
    n = N - 4
    db = torch.empty(F, n, device="cuda")
    cols = torch.empty(F, 200, device="cuda")
    stats = torch.empty(F, 4, device="cuda")
    mom = torch.empty(F, 4, device="cuda", dtype=torch.float64)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ctx.psd_dev(iq, N, F, db=db, window="hamming", epilogue=True, cols=cols, W=200, stats=stats, moments=mom)
        torch.cuda.synchronize()
    finally:
        ctx.set_stream(None)
    for name, t in (("db", db), ("cols", cols), ("stats", stats), ("moments", mom)):
        first = t[:K]
        full = first.repeat((F + K - 1) // K, 1)[:F]
        assert torch.equal(t, full), name
    got = db[:K].cpu().numpy().astype(np.float64)
    for f in range(K):
        want = O.psd_epilogue(O.psd_db(base[f]))
        assert np.max(np.abs(got[f] - want)) <= 1e-4, f


@pytest.mark.parametrize("mode", ["WFM", "NFM"])
def test_demod_full_bench_batch(ctx, mode):
    N, F, K, fs = 32768, 4096, 5, 2.4e6
    base = np.stack([synth.make("wbfm", N, seed=200 + i) for i in range(K)])
    iq = _dev(base).repeat((F + K - 1) // K, 1, 1)[:F].contiguous()
    plan = ctx.demod_plan(mode, fs, N)
    audio = torch.empty(F, plan.out_len, plan.channels, device="cuda")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ctx.demod_dev(plan, iq, F, audio)
        torch.cuda.synchronize()
    finally:
        ctx.set_stream(None)
    full = audio[:K].repeat((F + K - 1) // K, 1, 1)[:F]
    assert torch.equal(audio, full)
    got = audio[:K].cpu().numpy().astype(np.float64)
    for f in range(K):
        ref = O.demod(base[f], fs, mode)
        assert np.sqrt(np.mean((got[f] - ref) ** 2)) <= 1e-5, (mode, f)
