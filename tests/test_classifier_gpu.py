"""GPU: signal classifier against the reference run with its missing `welch` import
supplied (tests/golden/classifier.npz, oracle/make_golden_classifier.py) and against the oracle."""
import numpy as np
import pytest

from oracle import ref_dsp as O
from oracle.make_golden_classifier import CASES, make_case

pytestmark = pytest.mark.gpu

# tolerances: the Welch PSD is float32 in scipy for complex64 input (1e-6 relative); the modulation index
# divides two float32 variances, and np.unwrap accumulates its 2*pi corrections in float32, so a strong
# off-centre carrier (phase ramp of 1e4 rad) carries ~1e-3 relative rounding noise in the reference itself
TOL_FLAT = 1e-4
TOL_MI = 1e-2


@pytest.mark.parametrize("case", range(len(CASES)))
def test_classifier_golden(ctx, golden, case):
    g = golden("classifier")
    name, kw, n, fs = CASES[case]
    x = make_case(kw, n, seed=20 + case)
    labels, feat = ctx.classify(x, fs)
    want = g[name + "_feat"]
    assert labels[0] == str(g[name + "_label"]), (name, labels, feat, want)
    assert abs(feat[0, 0] - want[0]) <= 1e-6 * max(1.0, abs(want[0])), (name, feat[0, 0], want[0])
    assert abs(feat[0, 1] - want[1]) <= TOL_MI * abs(want[1]), (name, feat[0, 1], want[1])
    assert abs(feat[0, 2] - want[2]) <= TOL_FLAT * abs(want[2]), (name, feat[0, 2], want[2])


def test_classifier_batch_matches_oracle(ctx):
    fs = 2.4e6
    kinds = ["wbfm", "noise", "tone40", "halfband", "am", "ssb"]
    from pyspecsdr_b200 import synth
    x = np.stack([synth.make(k, 8192, seed=70 + i) for i, k in enumerate(kinds)])
    labels, feat = ctx.classify(x, fs)
    for i in range(len(x)):
        bw, mi, fl = O.classify_features(x[i], fs)
        assert labels[i] == O.classify_label(bw, mi, fl), (kinds[i], labels[i], feat[i], (bw, mi, fl))
        assert abs(feat[i, 0] - bw) <= 1e-6 * max(1.0, abs(bw))
        assert abs(feat[i, 1] - mi) <= TOL_MI * abs(mi)
        assert abs(feat[i, 2] - fl) <= TOL_FLAT * abs(fl)


def test_classify_signal_is_opt_in(ctx, monkeypatch):
    from pyspecsdr_b200 import signal_processing as sp, synth
    x = synth.make("noise", 4096, seed=1)
    monkeypatch.delenv("PSS_CLASSIFIER", raising=False)
    monkeypatch.setattr(sp, "ENABLE_CLASSIFIER", False)
    with pytest.raises(NameError):                 # the reference's behaviour (signal_processing.py:299)
        sp.classify_signal(x, 2.4e6, 0)
    monkeypatch.setattr(sp, "ENABLE_CLASSIFIER", True)
    assert sp.classify_signal(x, 2.4e6, 0) == O.classify_signal(x, 2.4e6)
    from pyspecsdr_b200.core import PssError
    with pytest.raises(PssError):                  # shorter than one Welch segment: unsupported, no fallback
        ctx.classify(x[:512], 2.4e6)
