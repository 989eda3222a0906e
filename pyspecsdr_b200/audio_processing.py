"""Drop-in for the numeric line of PySpecSDR's `audio_processing` module.

Only `write_audio_samples` contains arithmetic (audio_processing.py:34-38); the PortAudio probe and
the WAV open/close are host I/O and stay as they are in the reference (out of scope).
"""
import wave

import numpy as np

from . import signal_processing as _sp
from .filters import AUDIO_RATE as DEFAULT_SAMPLE_RATE


def start_audio_recording(filename, sample_rate=DEFAULT_SAMPLE_RATE):
    """audio_processing.py:24-31 (unchanged host I/O)."""
    wav_file = wave.open(filename, 'wb')
    wav_file.setnchannels(2)
    wav_file.setsampwidth(2)
    wav_file.setframerate(sample_rate)
    return wav_file


def write_audio_samples(wav_file, samples):
    """audio_processing.py:34-38: the float -> int16 pack runs on the GPU."""
    wav_file.writeframes(_sp._ctx().to_int16(np.asarray(samples)).tobytes())


def stop_audio_recording(wav_file):
    """audio_processing.py:41-43."""
    wav_file.close()
