"""GPU: the drop-in `signal_processing` module — same names, shapes and dtypes as the reference,
values within the north-star tolerances of the oracle."""
import numpy as np
import pytest

from oracle import ref_dsp as O
from pyspecsdr_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    from pyspecsdr_b200 import signal_processing
    return signal_processing


def test_exports_match_reference_namespace(sp):
    for name in ["compute_fft", "demodulate_signal", "demodulate_nfm", "demodulate_wfm", "demodulate_am",
                 "demodulate_ssb", "iq_correction", "mono_to_stereo", "bandpass_filter", "measure_signal_power",
                 "classify_signal", "np", "butter", "lfilter", "firwin", "hilbert", "decimate", "bilinear",
                 "resample_poly", "DEFAULT_SAMPLE_RATE", "BUTTER_ORDER"]:
        assert hasattr(sp, name), name
    assert sp.DEFAULT_SAMPLE_RATE == 22050 and sp.BUTTER_ORDER == 5


@pytest.mark.parametrize("n", [4096, 8192, 32768])      # 32768 = the app's default read, pyspecsdr.py:105,2236
def test_compute_fft(sp, n):
    x = synth.make("tone60", n, seed=1)
    y = sp.compute_fft(x)
    assert y.dtype == np.float64 and y.shape == (n,)
    assert np.max(np.abs(y - O.psd_db(x))) <= 1e-4


@pytest.mark.parametrize("mode,kind,n,fs", [("NFM", "wbfm", 32768, 2.4e6), ("WFM", "wbfm", 32768, 2.4e6),
                                            ("AM", "am", 32768, 1e6), ("USB", "ssb", 32768, 1e6),
                                            ("LSB", "ssb", 32768, 1e6), ("RAW", "tone40", 32768, 1e6),
                                            ("???", "noise", 256, 1e6)])
def test_demodulate_signal(sp, mode, kind, n, fs):
    x = synth.make(kind, n, seed=3)
    got = sp.demodulate_signal(x, fs, mode)
    ref = O.demod(x, fs, mode)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    scale = max(1.0, float(np.max(np.abs(ref)))) if mode == "RAW" else 1.0
    assert np.sqrt(np.mean((got - ref) ** 2)) <= 1e-5 * scale


def test_helpers(sp):
    x = synth.make("tone40", 4096, seed=2) * np.complex64(0.8 + 0.1j) + np.complex64(0.05 - 0.02j)
    got, ref = sp.iq_correction(x), O.iq_correct(x)
    assert got.dtype == ref.dtype == np.complex64
    assert np.sqrt(np.mean(np.abs(got - ref) ** 2)) <= 1e-5 * np.max(np.abs(ref))
    x = synth.make("am", 8192, seed=2)
    assert abs(sp.measure_signal_power(x) - O.signal_power_db(x)) <= 1e-4
    d = np.abs(synth.make("noise", 4096, seed=9)).astype(np.float32)
    for lo, hi, fs in ((0, 15000, 2.4e6), (300.0, 3000.0, 22050)):
        got, ref = sp.bandpass_filter(d, lo, hi, fs), O.bandpass(d, lo, hi, fs)
        assert got.dtype == np.float64 and got.shape == ref.shape
        assert np.sqrt(np.mean((got - ref) ** 2)) <= 1e-5 * np.max(np.abs(ref))
    np.testing.assert_array_equal(sp.mono_to_stereo(np.arange(5.0)), O.stereo(np.arange(5.0)))
    with pytest.raises(NameError):
        sp.classify_signal(x, 1e6, 1e4)


def test_int16_pack(sp, golden):
    """write_audio_samples' np.int16(samples * 32767) on the float64 array the app holds: bit-exact."""
    from pyspecsdr_b200 import audio_processing as ap
    g = golden("int16")
    pcm = ap.pcm16(g["audio"])
    assert pcm.dtype == np.int16 and pcm.shape == g["pcm"].shape
    np.testing.assert_array_equal(pcm, g["pcm"])
    # end to end through the WAV sink with the demodulator's own (float64-widened) output
    import io
    import wave
    x = synth.make("wbfm", 32768, seed=1)
    audio = sp.demodulate_signal(x, 2.4e6, "NFM")
    buf = io.BytesIO()
    w = ap.start_audio_recording(buf)
    ap.write_audio_samples(w, audio)
    ap.stop_audio_recording(w)
    buf.seek(0)
    with wave.open(buf, "rb") as r:
        assert (r.getnchannels(), r.getsampwidth(), r.getframerate()) == (2, 2, 22050)
        raw = r.readframes(r.getnframes())
    np.testing.assert_array_equal(np.frombuffer(raw, np.int16).reshape(-1, 2), np.int16(audio * 32767))


def test_audio_processing_namespace():
    from pyspecsdr_b200 import audio_processing as ap
    for name in ("init_audio_device", "start_audio_recording", "write_audio_samples", "stop_audio_recording", "sd",
                 "wave", "np", "DEFAULT_SAMPLE_RATE", "DEFAULT_BLOCK_SIZE"):
        assert hasattr(ap, name), name
    assert ap.DEFAULT_BLOCK_SIZE == 2048


def test_demodulate_wfm_direct_call_does_not_correct(sp):
    """The reference's demodulate_wfm (signal_processing.py:119-176) runs on the samples as given; only the
    dispatcher applies iq_correction (:222-225).  Direct call == oracle's demod_wfm without correction;
    dispatcher == correction + demod_wfm; and the two differ on an IQ-imbalanced input."""
    x = synth.make("wbfm", 32768, seed=4)
    x = (x.real * 1.2 + 1j * (x.imag * 0.5 + 0.3 * x.real)).astype(np.complex64)
    direct = sp.demodulate_wfm(x, 2.4e6)
    ref_direct = O.demod_wfm(x, 2.4e6)
    assert np.sqrt(np.mean((direct - ref_direct) ** 2)) <= 1e-5
    disp = sp.demodulate_signal(x, 2.4e6, "WFM")
    assert np.sqrt(np.mean((disp - O.demod(x, 2.4e6, "WFM")) ** 2)) <= 1e-5
    assert np.sqrt(np.mean((disp - direct) ** 2)) > 5e-4          # the oracle's two forms differ by 1.35e-3 here
    # corrected input through the direct call == the dispatcher, like the reference's two-step form
    two_step = sp.demodulate_wfm(sp.iq_correction(x), 2.4e6)
    assert np.sqrt(np.mean((two_step - disp) ** 2)) <= 1e-5
