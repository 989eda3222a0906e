"""GPU: edge fixtures SURVEY.md 8d lists for tests only — NaN propagation, all-zero blocks, error paths."""
import warnings

import numpy as np
import pytest

from oracle import ref_dsp as O
from pyspecsdr_b200 import synth
from pyspecsdr_b200.core import PssError

pytestmark = pytest.mark.gpu


def test_nan_sample_propagates_like_numpy(ctx):
    x = np.stack([synth.make("tone40", 1024, seed=s) for s in range(4)])
    x[1, 17] = np.nan + 0j                     # frame 1 shares its CTA with three clean frames
    got = ctx.psd(x, window="hamming")["db"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = O.psd_db(x)
    assert np.all(np.isnan(got[1])) and np.all(np.isnan(want[1]))
    for f in (0, 2, 3):
        assert np.max(np.abs(got[f] - want[f])) <= 1e-4
    res = ctx.psd(x, epilogue=True, want_stats=True)
    assert np.all(np.isnan(res["db"][1])) and np.isnan(res["stats"][1, 0]) and np.isnan(res["stats"][1, 1])
    for f in (0, 2, 3):
        assert np.max(np.abs(res["db"][f] - O.psd_epilogue(want[f]))) <= 1e-4


def test_all_zero_block_gives_nan_audio_like_the_reference(ctx):
    # the app never demodulates an all-zero read (pyspecsdr.py:2237); the reference would divide 0 by 0
    z = np.zeros((1, 8192), np.complex64)
    for mode, fs in (("NFM", 1.024e6), ("AM", 1e6), ("USB", 1e6)):
        got = ctx.demod(z, fs, mode)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = O.demod(z[0], fs, mode)
        assert np.all(np.isnan(ref)) and np.all(np.isnan(got)), mode


def test_single_sample_and_tiny_batches(ctx):
    x = synth.impulse(64, 5)
    assert np.max(np.abs(ctx.psd(x)["db"][0] - O.psd_db(x))) <= 1e-4
    assert ctx.demod(np.zeros((0, 8192), np.complex64), 1.024e6, "NFM").shape[0] == 0


def test_error_paths_raise_instead_of_falling_back(ctx):
    with pytest.raises(PssError):                          # not a power of two
        ctx.psd(synth.make("noise", 1000, seed=0))
    with pytest.raises(PssError):                          # beyond the four-step path
        ctx.psd(np.zeros(1 << 21, np.complex64))
    with pytest.raises(ValueError):                        # sample rate too low for a decimator
        ctx.demod(synth.make("noise", 4096, seed=0), 30e3, "NFM")
    with pytest.raises(ValueError):
        ctx.demod(synth.make("noise", 4096, seed=0), 1e6, "FSK")


def test_two_contexts_and_stream_adoption(ctx):
    import torch
    from pyspecsdr_b200 import core
    other = core.Context(0)
    try:
        x = np.stack([synth.make("wbfm", 4096, seed=s) for s in range(3)])
        a = ctx.psd(x)["db"]
        b = other.psd(x)["db"]
        np.testing.assert_array_equal(a, b)
        s = torch.cuda.Stream()
        other.set_stream(s.cuda_stream)
        iq = torch.from_numpy(x.view(np.float32).reshape(3, 4096, 2)).cuda()
        out = torch.empty(3, 4096, device="cuda")
        with torch.cuda.stream(s):
            other.psd_dev(iq, 4096, 3, db=out)
        s.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy(), a)
    finally:
        other.close()
