"""Drop-in for the arithmetic of PySpecSDR's `audio_processing` module.

The only numeric line there is the float -> int16 pack inside `write_audio_samples`
(audio_processing.py:36-38, repeated in io_manager.py:25-26); it runs on the GPU here
(`pss_audio_to_int16`).  The PortAudio probe (`init_audio_device`) is hardware I/O and is not provided:
keep importing it from the reference module.  The WAV container handling below exists only so that
`write_audio_samples(wav_file, samples)` can be exercised end to end.
"""
from __future__ import annotations

import wave

import numpy as np

from . import signal_processing as _dsp
from .filters import AUDIO_RATE as DEFAULT_SAMPLE_RATE

_STEREO, _PCM16_BYTES = 2, 2


def pcm16(samples) -> np.ndarray:
    """np.int16(samples * 32767) with C truncation, computed by the int16 kernel."""
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
