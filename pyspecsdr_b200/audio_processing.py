"""Drop-in for PySpecSDR's `audio_processing` module (`from audio_processing import *`, pyspecsdr.py:99).

The only numeric line there is the float -> int16 pack inside `write_audio_samples`
(audio_processing.py:36-38, repeated in io_manager.py:25-26); it runs on the GPU here
(`pss_audio_to_int16_f64` on the float64 array the caller holds: bit-identical to `np.int16(x * 32767)`).
Everything else in the module is device / file plumbing and is passed through with the reference's names so
the star import binds the same namespace: `init_audio_device` (a PortAudio probe, audio_processing.py:8-22),
`start_audio_recording`, `stop_audio_recording`, `sd`, `wave`, `np`, `DEFAULT_SAMPLE_RATE`,
`DEFAULT_BLOCK_SIZE`.
"""
from __future__ import annotations

import wave

import numpy as np

from . import signal_processing as _dsp
from .filters import AUDIO_RATE as DEFAULT_SAMPLE_RATE

DEFAULT_BLOCK_SIZE = 2048          # pyspecconst.py:4

try:                               # the reference imports it unconditionally (audio_processing.py:2)
    import sounddevice as sd
except (ImportError, OSError) as _e:      # no PortAudio on this host: the name exists, using it raises
    class _NoSoundDevice:
        PortAudioError = OSError
        _why = _e

        def __getattr__(self, name):
            raise ImportError(f"sounddevice is not available on this host ({self._why})")

    sd = _NoSoundDevice()

_STEREO, _PCM16_BYTES = 2, 2


def init_audio_device():
    """audio_processing.py:8-22: can a stereo float32 output stream be opened at the audio rate?"""
    try:
        probe = sd.OutputStream(channels=_STEREO, samplerate=DEFAULT_SAMPLE_RATE, blocksize=DEFAULT_BLOCK_SIZE,
                                dtype=np.float32)
        probe.close()
        return True
    except sd.PortAudioError as e:
        print(f"Audio initialization error: {e}")
        return False


def pcm16(samples) -> np.ndarray:
    """np.int16(samples * 32767) with C truncation, computed by the int16 kernel (fp64 in, like numpy)."""
    return _dsp._ctx().to_int16(np.asarray(samples))


def write_audio_samples(wav_file, samples):
    """Same call as the reference: append one block of float audio to an open 16-bit WAV."""
    wav_file.writeframes(pcm16(samples).tobytes())


def start_audio_recording(filename, sample_rate=DEFAULT_SAMPLE_RATE):
    sink = wave.open(filename, "wb")
    sink.setparams((_STEREO, _PCM16_BYTES, int(sample_rate), 0, "NONE", "not compressed"))
    return sink


def stop_audio_recording(wav_file):
    wav_file.close()
