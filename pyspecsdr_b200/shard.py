"""Frame sharding across GPUs (one process per GPU) and the scanner's wide-band gather.

Every block of the hot path is independent (zero filter state, per-block normalisation:
signal_processing.py:108, 115, 126-160, 191-194, 204-216), so a batch or capture file shards by
contiguous frame ranges with NO data-path collective.  The only exchange step is the scanner
(pyspecsdr.py:2514-2590 / 1022-1093): each rank sweeps its share of the frequency steps and the
per-step (peak, count[, dB row]) records are all-gathered so every rank holds the stitched sweep.
Works with any torch.distributed backend: NCCL on device tensors, gloo on CPU tensors (tests).
"""
from __future__ import annotations

from typing import Optional, Tuple


def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of frames owned by `rank`: r*F//R .. (r+1)*F//R."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return rank * n_frames // world, (rank + 1) * n_frames // world


def shard_sizes(n_frames: int, world: int):
    return [frame_range(n_frames, r, world)[1] - frame_range(n_frames, r, world)[0] for r in range(world)]


def gather_sweep(peak, count, rows=None, n_steps: Optional[int] = None, group=None):
    """All-gather the per-rank scanner results into step order.

    peak [n_local] float32, count [n_local] int32, rows [n_local, N] float32 or None are this rank's
    results for its `frame_range` of the `n_steps` sweep (torch tensors on the backend's device).
    Returns (peak [n_steps], count [n_steps], rows [n_steps, N] | None), identical on every rank and
    bitwise equal to what a single rank computing every step produces.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return peak, count, rows
    world = dist.get_world_size(group)
    if n_steps is None:
        t = torch.tensor([peak.shape[0]], device=peak.device, dtype=torch.int64)
        dist.all_reduce(t, group=group)
        n_steps = int(t.item())
    sizes = shard_sizes(n_steps, world)
    m = max(sizes)
    if min(sizes) == m:
        # equal shares (1000 steps over 1 / 2 / 4 / 8 ranks): every rank's records land at their final place in
        # one all_gather_into_tensor each - the (peak, count) pairs travel together as 8-byte records
        rec = torch.empty(m, 2, dtype=torch.int32, device=peak.device)
        rec[:, 0] = peak.view(torch.int32)
        rec[:, 1] = count
        allrec = torch.empty(n_steps, 2, dtype=torch.int32, device=peak.device)
        dist.all_gather_into_tensor(allrec, rec, group=group)
        allrows = None
        if rows is not None:
            allrows = torch.empty((n_steps,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
            dist.all_gather_into_tensor(allrows, rows.contiguous(), group=group)
        return allrec[:, 0].contiguous().view(torch.float32), allrec[:, 1].contiguous(), allrows

    def ag(x):
        pad = torch.zeros((m,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[: x.shape[0]] = x
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad, group=group)
        return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)

    return ag(peak), ag(count), (ag(rows) if rows is not None else None)


def sweep(ctx, frames_dev, N: int, n_steps: int, rank: int, world: int, want_rows: bool = False, rel_db: float = 20.0,
          group=None):
    """Scanner sweep sharded over ranks: `frames_dev` holds THIS rank's steps ([n_local, N] complex64 as
    a torch CUDA tensor of float pairs).  Returns the stitched (peak, count, rows) on every rank."""
    import torch

    lo, hi = frame_range(n_steps, rank, world)
    n_local = hi - lo
    dev = frames_dev.device
    peak = torch.empty(n_local, device=dev, dtype=torch.float32)
    count = torch.empty(n_local, device=dev, dtype=torch.int32)
    rows = torch.empty(n_local, N, device=dev, dtype=torch.float32) if want_rows else None
    if n_local:
        # Stream contract of the *_dev calls (include/pss.h): the scan runs on the CONTEXT's stream, which has
        # no ordering with torch's unless the caller adopted it with ctx.set_stream().  So: finish whatever
        # produced `frames_dev` on torch's current stream, enqueue the scan, and drain the context's stream
        # before the collective (which runs on torch's / NCCL's streams) may read peak / count / rows.
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()
        ctx.scan_dev(frames_dev, N, n_local, peak, count, rows=rows, rel_db=rel_db)
        ctx.sync()
    return gather_sweep(peak, count, rows, n_steps, group)
