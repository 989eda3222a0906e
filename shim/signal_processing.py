"""Put this directory ahead of the PySpecSDR tree on sys.path: `from signal_processing import *`
(pyspecsdr.py:98) and `from signal_processing import bandpass_filter` (decoders.py:3) then bind the
B200 implementations, with zero edits to the application."""
from pyspecsdr_b200.signal_processing import *          # noqa: F401,F403
