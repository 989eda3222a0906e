"""Shared by the CPU (oracle) and GPU display tests: the seeded dB rows behind tests/golden/display.npz and
the reconstruction of what the reference's draw_* functions put on the screen from the integer planes.
The cell logic restates the drawing loops only (which cell, which glyph, which attribute); every number
in a plane comes from the implementation under test."""
import hashlib

import numpy as np

from oracle import ref_dsp as O
from pyspecsdr_b200 import synth

WATERFALL_GLYPHS = np.array([ord(c) for c in ".-=#"])
GRADIENT_GLYPHS = np.array([ord(c) for c in " ._-=+*#@"])


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def golden_rows(g):
    """The 34 fp64 dB rows the golden display cells were drawn from (oracle/make_golden.py); the checksum
    proves they are bit-identical to the reference's own rows."""
    rows = []
    for s in range(34):
        x = synth.make("wbfm" if s % 2 else "tone40", 4096, seed=100 + s)
        rows.append(O.psd_epilogue(O.psd_db(x)))
    assert digest(np.array(rows)) == str(g["rows_in"])
    return rows


def waterfall_cells(level, colour):
    """(char, attr) planes as FakeScreen.cells rasterises draw_waterfall (pyspecsdr.py:1390-1400)."""
    return WATERFALL_GLYPHS[level], (10 + colour.astype(np.int64)) << 8


def gradient_cells(chars, colour):
    """draw_gradient_waterfall, pyspecsdr.py:1691-1700."""
    return GRADIENT_GLYPHS[chars], (10 + colour.astype(np.int64)) << 8


def persistence_stars(ys, colours, H):
    """The '*' draws of draw_persistence in call order (pyspecsdr.py:1543-1560): traces oldest first,
    columns left to right, only rows 0 <= y < H.  `ys` [traces, W] int, `colours` [traces]."""
    return np.array([(int(y) + 2, x + 8, int(c) << 8) for yrow, c in zip(ys, colours)
                     for x, y in enumerate(yrow) if 0 <= y < H], dtype=np.int64)


def surface_cells(mag, H, Wt):
    """Cells draw_surface_plot touches with '#' (projection loop, pyspecsdr.py:1594-1601)."""
    ang = np.radians(O.SURFACE_ANGLE)
    cells = set()
    for x, m in enumerate(mag):
        for y in range(int(m)):
            sx = int(x - y * np.cos(ang)) + 8
            sy = int(H - 2 - y * np.sin(ang))
            if 0 <= sx < Wt and 2 <= sy < H - 1:
                cells.add((sy, sx))
    return cells


def spectrum_cells(cols, dh):
    """Last (char, attr) per cell of draw_spectrogram's bar loop (pyspecsdr.py:455-493)."""
    import curses
    out = {}
    for x, v in enumerate(cols):
        height = min(int(v * dh), dh)
        for y in range(dh):
            if y < dh - height:
                out[(y + 2, x + 7)] = (ord(" "), 1 << 8)
                continue
            rel = (y - (dh - height)) / height if height > 0 else 0
            if v > 0.8:
                ch, col = ("#" if rel > 0.5 else "="), 14
            elif v > 0.4:
                ch, col = ("=" if rel > 0.5 else "-"), 13
            elif v > 0.2:
                ch, col = ("-" if rel > 0.5 else "."), 12
            elif rel > 0.7:
                ch, col = ".", 11
            else:
                ch, col = " ", 10
            out[(y + 2, x + 7)] = (ord(ch), (col << 8) | curses.A_BOLD)
    return out


def assert_equal_or_on_boundary(got, ref_value, scale, tol, what=""):
    """Integer planes computed from float32 spectra: `got` must equal int(ref_value * scale) except where
    the reference's own value sits within `tol` (in units of the normalised value) of a quantisation
    boundary - the only place where a spectrum that is within the dB tolerance may legitimately land in
    the neighbouring cell.  Returns the number of such boundary cells."""
    ref_value = np.asarray(ref_value, dtype=np.float64)
    want = (ref_value * scale).astype(np.int64)
    bad = np.asarray(got).astype(np.int64) != want
    if not bad.any():
        return 0
    dist = np.abs(ref_value[bad] * scale - np.round(ref_value[bad] * scale))
    assert np.all(dist <= tol * scale), f"{what}: {int(bad.sum())} mismatches, worst boundary distance {dist.max():.3e}"
    assert np.all(np.abs(np.asarray(got).astype(np.int64)[bad] - want[bad]) == 1), what
    return int(bad.sum())
