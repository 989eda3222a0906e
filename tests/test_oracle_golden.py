"""CPU: the oracle (oracle/ref_dsp.py) against golden vectors produced by the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle; bit-exact unless stated."""
import hashlib

import numpy as np
import pytest

import _display_cells as D
from oracle import ref_dsp as O
from pyspecsdr_b200 import synth


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_versions_match_golden(golden):
    import scipy
    g = golden("psd")
    # the goldens are tied to the library versions they were produced with
    assert str(g["_numpy"]) == np.__version__
    assert str(g["_scipy"]) == scipy.__version__


@pytest.mark.parametrize("kind,n", [(k, n) for k in ("noise", "tone40", "tone60", "wbfm")
                                    for n in (1024, 4096, 8192, 16384)
                                    if n <= 4096 or k in ("tone60", "wbfm")])
def test_psd(golden, kind, n):
    g = golden("psd")
    x = synth.make(kind, n, seed=n % 97)
    assert digest(x) == str(g[f"{kind}_{n}_in"])
    np.testing.assert_array_equal(O.psd_db(x), g[f"{kind}_{n}"])


def test_psd_edge(golden):
    g = golden("psd")
    np.testing.assert_array_equal(O.psd_db(synth.impulse(1024, 3)), g["impulse_1024"])
    np.testing.assert_array_equal(O.psd_db(np.zeros(1024, np.complex64)), g["zeros_1024"])
    assert np.all(g["zeros_1024"] == -100.0)          # 10*log10(1e-10)


def test_psd_batched_equals_rowwise():
    x = np.stack([synth.make("tone40", 1024, seed=s) for s in range(3)])
    np.testing.assert_array_equal(O.psd_db(x), np.stack([O.psd_db(r) for r in x]))


@pytest.mark.parametrize("kind,n", [("noise", 1024), ("tone40", 4096), ("wbfm", 4096), ("tone60", 8192),
                                    ("halfband", 4096), ("halfband", 1024)])
def test_epilogue(golden, kind, n):
    g = golden("epilogue")
    x = synth.make(kind, n, seed=5)
    assert digest(x) == str(g[f"{kind}_{n}_in"])
    out = O.psd_epilogue(O.psd_db(x))
    assert out.shape == (n - 4,)
    np.testing.assert_array_equal(out, g[f"{kind}_{n}"])


def test_scanner(golden):
    g = golden("scanner")
    frames = synth.scanner_frames(24, 2048, seed=3)
    assert digest(frames) == str(g["in"])
    res = [O.scan_step(f, 2.4e6) for f in frames]
    np.testing.assert_array_equal([r[0] for r in res], g["peak"])
    np.testing.assert_array_equal([r[1] for r in res], g["count"])
    np.testing.assert_array_equal([r[2] for r in res], g["bandwidth"])
    f8 = synth.scanner_frames(6, 8192, seed=4)
    res = [O.scan_step(f, 2.4e6) for f in f8]
    np.testing.assert_array_equal([r[0] for r in res], g["peak8k"])
    np.testing.assert_array_equal([r[1] for r in res], g["count8k"])


DEMOD_CASES = [
    ("NFM", "wbfm", 32768, 2.4e6), ("NFM", "noise", 32768, 2.4e6), ("NFM", "wbfm", 16385, 1.024e6),
    ("WFM", "wbfm", 32768, 2.4e6), ("WFM", "noise", 32768, 2.4e6), ("WFM", "wbfm", 16385, 1.024e6),
    ("AM", "am", 8192, 1e6), ("AM", "noise", 8192, 1e6),
    ("USB", "ssb", 8192, 1e6), ("LSB", "ssb", 8192, 1e6), ("USB", "noise", 4097, 1e6),
    ("RAW", "tone40", 4096, 1e6), ("XXX", "noise", 256, 1e6),
]


@pytest.mark.parametrize("mode,kind,n,fs", DEMOD_CASES)
def test_demod(golden, mode, kind, n, fs):
    g = golden("demod")
    key = f"{mode}_{kind}_{n}_{int(fs)}"
    x = synth.make(kind, n, seed=11)
    assert digest(x) == str(g[key + "_in"])
    y = O.demod(x, fs, mode)
    assert tuple(g[key + "_shape"]) == y.shape
    assert str(g[key + "_dtype"]) == str(y.dtype)
    want = g[key]
    got = y[:, 0] if (y.ndim == 2 and mode != "WFM") else y
    np.testing.assert_array_equal(got, want)
    if y.ndim == 2:
        # reference quirk: every stereo output has (numerically) identical channels; for WFM the
        # pilot path is sin(0|pi) = 0|1.2e-16, so L-R is ~1e-12 of full scale
        assert float(g[key + "_lr_maxdiff"]) <= 1e-10
        assert np.max(np.abs(y[:, 0] - y[:, 1])) <= 1e-10


def test_usb_equals_lsb(golden):
    g = golden("demod")
    np.testing.assert_array_equal(g["USB_ssb_8192_1000000"], g["LSB_ssb_8192_1000000"])


def test_helpers(golden):
    g = golden("demod")
    x = synth.make("tone40", 4096, seed=2) * np.complex64(0.8 + 0.1j) + np.complex64(0.05 - 0.02j)
    assert digest(x) == str(g["iqcorr_in"])
    np.testing.assert_array_equal(O.iq_correct(x), g["iqcorr"])
    x = synth.make("am", 8192, seed=2)
    assert O.signal_power_db(x) == float(g["power_db"])
    d = np.abs(synth.make("noise", 4096, seed=9)).astype(np.float32)
    np.testing.assert_array_equal(O.bandpass(d, 0, 15000, 2.4e6), g["bandpass_lp"])
    np.testing.assert_array_equal(O.bandpass(d, 300.0, 3000.0, 22050), g["bandpass_bp"])
    np.testing.assert_array_equal(O.stereo(np.arange(5.0)), g["stereo"])


def test_int16(golden):
    g = golden("int16")
    np.testing.assert_array_equal(O.to_int16(g["audio"]), g["pcm"])


# ------------------------------------------------------------------ display math, pinned through
# what the reference's draw_* functions actually drew on a recording fake screen
def _rows(golden):
    g = golden("display")
    return g, D.golden_rows(g)


def test_waterfall_display(golden):
    g, rows = _rows(golden)
    W = int(g["W"]) - 8
    hist = []
    for s, r in enumerate(rows):
        norm, (lo, hi), colour, level = O.waterfall_accumulate(hist, r, W)
        if s in (0, 5, 33):
            ch, at = D.waterfall_cells(level, colour)
            np.testing.assert_array_equal(ch, g[f"waterfall_{s}_char"])
            np.testing.assert_array_equal(at, g[f"waterfall_{s}_attr"])
    assert len(hist) == 30


def test_gradient_display(golden):
    g, rows = _rows(golden)
    W = int(g["W"]) - 10
    hist = []
    for s, r in enumerate(rows):
        norm, _, chars, colour = O.gradient_accumulate(hist, r, W)
        if s in (0, 33):
            ch, at = D.gradient_cells(chars, colour)
            np.testing.assert_array_equal(ch, g[f"gradient_{s}_char"])
            np.testing.assert_array_equal(at, g[f"gradient_{s}_attr"])


def test_persistence_display(golden):
    g, rows = _rows(golden)
    H, W = int(g["H"]) - 4, int(g["W"]) - 8
    hist = []
    for s, r in enumerate(rows[:14]):
        ys, colours, _ = O.persistence_accumulate(hist, r, W, H)
        if s in (0, 13):
            np.testing.assert_array_equal(D.persistence_stars(ys, colours, H), g[f"persistence_{s}_stars"])


def test_surface_display(golden):
    g, rows = _rows(golden)
    H, Wt = int(g["H"]), int(g["W"])
    mag, _ = O.surface_row(rows[3], Wt - 8)
    # the reference overdraws cells; compare the set of (y, x) it touched with '#'
    want = {(int(a), int(b)) for a, b, _ in g["surface_hash_cells"]}
    assert D.surface_cells(mag, H, Wt) == want


def test_spectrum_display(golden):
    g, rows = _rows(golden)
    H, Wt = int(g["H"]), int(g["W"])
    dh, dw = H - 4, Wt - 7
    cols, _ = O.spectrum_normalise(rows[3], dw)
    want = {(int(y), int(x)): (int(c), int(a)) for y, x, c, a in g["spectrum_cells"]}
    got = D.spectrum_cells(cols, dh)
    assert all(want[k] == v for k, v in got.items()) and len(got) == dh * dw


def test_classifier_oracle_matches_reference_golden(golden):
    """§8f-4: the oracle's classify_features / classify_label against the reference's own functions run
    with the missing `welch` import supplied (oracle/make_golden_classifier.py)."""
    from oracle.make_golden_classifier import CASES, make_case
    g = golden("classifier")
    for i, (name, kw, n, fs) in enumerate(CASES):
        x = make_case(kw, n, seed=20 + i)
        bw, mi, fl = O.classify_features(x, fs)
        want = g[name + "_feat"]
        assert bw == want[0] and mi == want[1] and fl == want[2], name
        assert O.classify_label(bw, mi, fl) == str(g[name + "_label"]), name
