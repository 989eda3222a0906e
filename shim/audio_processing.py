"""Shadows PySpecSDR's audio_processing module (`from audio_processing import *`, pyspecsdr.py:99)."""
from pyspecsdr_b200.audio_processing import *           # noqa: F401,F403
