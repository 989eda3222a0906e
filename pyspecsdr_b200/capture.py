"""IQ capture-file ingest: the on-disk format either side of the hot path.

The reference records a capture with `np.save(filename, samples)` and reads it back with
`np.load(filename)` (pyspecsdr.py:813-824): a `.npy` file holding one 1-D complex64 array.  This module
turns the kernels into an offline capture processor: the file is memory-mapped, cut into main-loop reads
("blocks" of `block` samples, like `sdr.read_samples((2**SAMPLES)*256)`, pyspecsdr.py:2236), the block
range is sharded over ranks (one process per GPU, no collective: blocks are independent), and each rank
streams its share through pinned staging buffers into `pss_pipeline_c64`.
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional, Tuple

import numpy as np

from . import shard


def open_capture(path: str) -> np.ndarray:
    """Memory-map a capture written by the reference's record_signal (complex64, 1-D)."""
    a = np.load(path, mmap_mode="r")
    if a.ndim != 1 or a.dtype != np.complex64:
        raise ValueError(f"{path}: expected the reference's capture format (1-D complex64), got "
                         f"{a.dtype} with shape {a.shape}")
    return a


def block_range(n_samples: int, block: int, rank: int = 0, world: int = 1) -> Tuple[int, int]:
    """Blocks [lo, hi) owned by `rank`.  A trailing partial block is dropped, like a read that
    returned fewer samples than requested is skipped by the app (pyspecsdr.py:2237-2242)."""
    return shard.frame_range(n_samples // block, rank, world)


def iter_chunks(capture: np.ndarray, block: int, lo: int, hi: int, chunk_blocks: int) -> Iterator[Tuple[int, np.ndarray]]:
    """(first block index, [n, block] view) pieces of the block range [lo, hi)."""
    for b0 in range(lo, hi, chunk_blocks):
        b1 = min(hi, b0 + chunk_blocks)
        yield b0, capture[b0 * block:b1 * block].reshape(b1 - b0, block)


def process_capture(ctx, path: str, fs: float, mode: str = "NFM", block: int = 32768, n_fft: int = 4096,
                    W: int = 200, rows_max: int = 30, rank: int = 0, world: int = 1,
                    chunk_blocks: int = 1024, limit_blocks: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Run the whole main-loop iteration (demodulate_signal + compute_fft/epilogue per `n_fft` frame +
    waterfall accumulate) over this rank's share of a capture file.

    Returns dict(first_block, audio [nb, out_len, ch], cols [nb*fpb, W], stats [nb*fpb, 4],
    norm [nb, rows_max, W], minmax [nb, 2]).  The waterfall history restarts at every chunk boundary's
    first block only if `chunk_blocks` is smaller than the rank's share AND the caller asked for it;
    here history is carried across chunks by re-feeding the last rows_max frames' columns, so the
    result does not depend on the chunking."""
    cap = open_capture(path)
    lo, hi = block_range(len(cap), block, rank, world)
    if limit_blocks is not None:
        hi = min(hi, lo + limit_blocks)
    nb = hi - lo
    fpb = block // n_fft
    plan = ctx.demod_plan(mode, fs, block) if mode else None
    out = {
        "first_block": lo,
        "cols": np.empty((nb * fpb, W), np.float32), "stats": np.empty((nb * fpb, 4), np.float32),
        "norm": np.empty((nb, rows_max, W), np.float32), "minmax": np.empty((nb, 2), np.float32),
    }
    if plan:
        out["audio"] = np.empty((nb, plan.out_len, plan.channels), np.float32)
    if nb == 0:
        return out
    stage = ctx.pinned_empty((min(chunk_blocks, nb), block), np.complex64)
    for b0, view in iter_chunks(cap, block, lo, hi, chunk_blocks):
        n = len(view)
        stage[:n] = view                                   # page cache / disk -> pinned staging
        res = ctx.pipeline(stage[:n], fs, mode, n_fft=n_fft, W=W, rows_max=rows_max)
        k = b0 - lo
        out["cols"][k * fpb:(k + n) * fpb] = res["cols"]
        out["stats"][k * fpb:(k + n) * fpb] = res["stats"]
        if plan:
            out["audio"][k:k + n] = res["audio"]
    # the display history is a function of cols/stats only: render it over the whole share so that
    # chunk boundaries do not cut the 30-row history
    norm, mm = ctx.display_render(out["cols"], out["stats"], rows_max=rows_max, first=fpb - 1, step=fpb, n_renders=nb)
    out["norm"], out["minmax"] = norm, mm
    return out
