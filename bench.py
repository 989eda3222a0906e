#!/usr/bin/env python
"""bench.py — IQ Msamples/s through FFT+PSD+waterfall+FM-demod (BASELINE.json configs[1], "C2").

A step = one pass of the hot path over one batch of synthetic IQ per GPU:
    for every 32768-sample read ("block") of a 2.4 MS/s WBFM-tone stream:
        demodulate_signal(block, fs, mode)                      (signal_processing.py:220-240)
        for each 4096-sample frame: compute_fft + smoothing + median clamp + peak/avg + W-column
        resample                                                (signal_processing.py:243-264,
                                                                 pyspecsdr.py:2278-2283, 388-389)
        waterfall history (30 rows) normalisation after the block   (pyspecsdr.py:1351-1398)
`value`  : device-resident inputs, CUDA-event timed, max over ranks (weak scaling: each rank owns
           its own blocks; no collective on the data path).
`e2e`    : the same work through the host-pointer C ABI (pss_pipeline_c64) from pinned host memory,
           copies inside the timed region.
`--impl reference` : the reference's CPU path (oracle port: identical numpy/scipy calls) on all host
           cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 2.4e6
N_BLOCK = 32768
N_FFT = 4096
W_COLS = 200
ROWS_MAX = 30
METRIC = "iq_msamples_per_s_fft_psd_waterfall_fm_demod"


# ----------------------------------------------------------------------------- synthetic input
def make_blocks_torch(n_blocks, seed, device):
    """2.4 MS/s WBFM tone (1 kHz tone, 75 kHz deviation) + noise at -40 dBc, complex64 as [..,2] f32."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(1234 + seed)
    t = torch.arange(N_BLOCK, device=device, dtype=torch.float64) / FS
    out = torch.empty(n_blocks, N_BLOCK, 2, device=device, dtype=torch.float32)
    chunk = 256
    for b0 in range(0, n_blocks, chunk):
        nb = min(chunk, n_blocks - b0)
        ph0 = torch.rand(nb, 1, device=device, dtype=torch.float64, generator=g) * 6.283185307179586
        m = torch.sin(2 * np.pi * 1e3 * t[None, :] + ph0)
        ph = 2 * np.pi * 75e3 * torch.cumsum(m, dim=1) / FS
        sig = torch.stack([torch.cos(ph), torch.sin(ph)], dim=-1)
        noise = torch.randn(nb, N_BLOCK, 2, device=device, dtype=torch.float64, generator=g) * (0.01 / np.sqrt(2))
        out[b0:b0 + nb] = (sig + noise).to(torch.float32)
    return out


def make_blocks_numpy(n_blocks, seed):
    from pyspecsdr_b200 import synth
    return np.stack([synth.wbfm(N_BLOCK, seed=seed * 100003 + b, fs=FS) for b in range(n_blocks)])


# ----------------------------------------------------------------------------- CPU reference arm
_CPU_BLOCKS = None      # set before the fork pool is created so workers inherit the data without pickling


def _oracle_block_worker(args):
    """Process a contiguous range of blocks exactly as the reference's main loop would."""
    lo, hi, mode = args
    from oracle import ref_dsp as O
    hist = []
    acc = 0.0
    for blk in _CPU_BLOCKS[lo:hi]:
        audio = O.demod(blk, FS, mode)
        acc += float(audio[0, 0])
        for f in range(N_BLOCK // N_FFT):
            row = O.psd_epilogue(O.psd_db(blk[f * N_FFT:(f + 1) * N_FFT]))
            pk, av = O.peak_avg(row)
            if f < N_BLOCK // N_FFT - 1:
                hist.append(row)
                if len(hist) > ROWS_MAX:
                    hist.pop(0)
        norm, _, _, _ = O.waterfall_accumulate(hist, row, W_COLS, ROWS_MAX)
        acc += float(norm[0, 0]) + pk + av
    return acc


class CpuReference:
    """The reference's CPU path (oracle port: the same numpy/scipy calls) over all host cores: one
    forked worker per core, disjoint contiguous block ranges, single-threaded BLAS/OpenMP per worker."""

    def __init__(self, blocks_c64, mode, cores=None):
        global _CPU_BLOCKS
        import multiprocessing as mp
        os.environ["OMP_NUM_THREADS"] = "1"
        os.environ["OPENBLAS_NUM_THREADS"] = "1"
        os.environ["MKL_NUM_THREADS"] = "1"
        _CPU_BLOCKS = blocks_c64
        self.mode = mode
        self.nb = len(blocks_c64)
        self.cores = max(1, min(cores or os.cpu_count() or 1, self.nb))
        self.pool = mp.get_context("fork").Pool(self.cores)
        self.parts = [(i * self.nb // self.cores, (i + 1) * self.nb // self.cores, mode) for i in range(self.cores)]
        self.pool.map(_oracle_block_worker, [(lo, min(hi, lo + 1), mode) for lo, hi, _ in self.parts])   # warm

    def step(self):
        """One pass over the sample; returns wall seconds."""
        t0 = time.perf_counter()
        self.pool.map(_oracle_block_worker, self.parts)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                mx = float(p[1])
                if t_begin - 0.05 <= ts <= t_end + 0.15:
                    sm.append(float(p[0]))
                    for nm, v in zip(names, p[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- NUMA placement
def bind_to_gpu_numa_node(device_index):
    """Pin this rank's CPU affinity to the NUMA node its GPU hangs off (sysfs), so the pinned staging
    buffers it allocates next are node-local and N ranks do not all pull their host->device copies
    across the inter-socket link.  Returns (original affinity, note); no-op when sysfs lacks the info."""
    try:
        import torch
        orig = os.sched_getaffinity(0)
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return orig, "numa_node=-1 (single node or not exposed)"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= orig
        if cpus:
            os.sched_setaffinity(0, cpus)
            return orig, f"bound to NUMA node {node} ({len(cpus)} cpus)"
        return orig, f"NUMA node {node} has no allowed cpus"
    except Exception as e:                                    # never fail the bench over placement
        return None, f"not bound ({type(e).__name__})"


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="WFM", choices=["WFM", "NFM"])
    ap.add_argument("--blocks", type=int, default=4096, help="32768-sample blocks per GPU per step (4096 = 1 GiB)")
    ap.add_argument("--cpu-blocks", type=int, default=1536, help="blocks in the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"C2: 4096-pt Hamming FFT PSD + 5-bin smoothing + median clamp + peak/avg + {W_COLS}-col resample, "
                f"30-row waterfall normalisation per block, {args.mode} demod on 32768-sample blocks, "
                f"2.4 MS/s synthetic WBFM tone (-40 dBc noise)")

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return
        args.cpu_blocks = min(args.cpu_blocks, 512)        # keeps K steps within a few minutes on any host
        blocks = make_blocks_numpy(args.cpu_blocks, seed=0)
        ref = CpuReference(blocks, args.mode)
        for _ in range(args.warmup):
            ref.step()
        walls = [ref.step() for _ in range(args.steps)]
        ref.close()
        ms = float(np.mean(walls)) * 1e3
        v = args.cpu_blocks * N_BLOCK / (ms * 1e-3) / 1e6
        info = {"cores": ref.cores}
        sample = f"{args.cpu_blocks} blocks x {N_BLOCK} samples per step, numpy/scipy oracle port, fork pool"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "mode": args.mode, "blocks_per_step": args.cpu_blocks},
            "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": info["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from pyspecsdr_b200 import core

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = core.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)

    nb = args.blocks
    fpb = N_BLOCK // N_FFT
    F = nb * fpb
    n_bins = N_FFT - 4
    iq = make_blocks_torch(nb, seed=rank, device=dev)                      # rank-private blocks (frame sharding)
    plan = ctx.demod_plan(args.mode, FS, N_BLOCK)
    db = torch.empty(F, n_bins, device=dev, dtype=torch.float32)
    cols = torch.empty(F, W_COLS, device=dev, dtype=torch.float32)
    stats = torch.empty(F, 4, device=dev, dtype=torch.float32)
    norm = torch.empty(nb, ROWS_MAX, W_COLS, device=dev, dtype=torch.float32)
    minmax = torch.empty(nb, 2, device=dev, dtype=torch.float32)
    audio = torch.empty(nb, plan.out_len, plan.channels, device=dev, dtype=torch.float32)
    moments = torch.empty(F, 4, device=dev, dtype=torch.float64)     # PSD by-product consumed by the WFM demod
    torch.cuda.synchronize()

    names = ["psd_kernel<12,f64,smooth>", "display_render_kernel", f"demod_decim_kernel<{args.mode}>"]

    def step(ev=None):
        if ev is not None:
            ev[0].record(stream)
        ctx.psd_dev(iq, N_FFT, F, db=db, window="hamming", epilogue=True, cols=cols, W=W_COLS, stats=stats,
                    moments=moments)
        if ev is not None:
            ev[1].record(stream)
        ctx.display_render_dev(cols, stats, W_COLS, F, norm, minmax, rows_max=ROWS_MAX, first=fpb - 1, step=fpb,
                               n_renders=nb)
        if ev is not None:
            ev[2].record(stream)
        ctx.demod_dev(plan, iq, nb, audio, moments=moments, frames_per_block=fpb)
        if ev is not None:
            ev[3].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        barrier()
        clocks = ClockSampler(local_rank) if rank == 0 else None
        if clocks:
            clocks.start()
            time.sleep(0.3)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
        l0 = ctx.launches
        barrier()
        t_begin = time.time()
        e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_start.record(stream)
        for k in range(args.steps):
            step(evs[k])
        e_stop.record(stream)
        barrier()
        t_end = time.time()
        launches = ctx.launches - l0
        total_ms = e_start.elapsed_time(e_stop)
    clock_info = clocks.stop(t_begin, t_end) if clocks else None

    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    ln = torch.tensor([launches], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    samples_per_step = world * nb * N_BLOCK
    value = samples_per_step / ms_per_step / 1e3                            # Msamples/s, whole job

    # per-kernel durations inside the timed region (rank 0) -> the dominant kernel's roofline
    kms = np.zeros(3)
    for ev in evs:
        for i in range(3):
            kms[i] += ev[i].elapsed_time(ev[i + 1])
    kms /= args.steps
    alg_bytes = [F * N_FFT * 8 + F * n_bins * 4 + F * W_COLS * 4 + F * 16,
                 nb * ROWS_MAX * W_COLS * 4 * 2,
                 nb * N_BLOCK * 8 + audio.numel() * 4]
    dom = int(np.argmax(kms))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    if os.path.exists(peaks_path):
        try:                                  # driver-written; tolerate a different key spelling
            mp = json.load(open(peaks_path))
            key = "hbm_gbs" if "hbm_gbs" in mp else next(k for k in mp if "hbm" in k.lower())
            # kernels here are timed inside a long step sequence: the sustained figure when both are given
            val = mp[key]
            if isinstance(val, dict):
                val = val.get("sustained", val.get("burst", val))
            val = float(val if not isinstance(val, dict) else next(iter(val.values())))
            if val < 100.0:                   # TB/s -> GB/s
                val *= 1000.0
            peak, peak_src = val, f"MEASURED_PEAKS.json {key} (of measured)"
        except Exception:
            pass
    achieved = alg_bytes[dom] / (kms[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.mode, {}).get(names[dom])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes[dom], "kernel_ms": float(kms[dom]),
                "all_kernels_ms": {n: float(m) for n, m in zip(names, kms)},
                "all_kernels_frac": {n: float(b / (m * 1e-3) / 1e9 / peak) for n, b, m in zip(names, alg_bytes, kms)}}

    # parity spot-check on this run's data (rank 0): first two blocks against the oracle
    parity = None
    cpu_base = None
    e2e = None
    if rank == 0:
        from oracle import ref_dsp as O
        xs = iq[:2].cpu().numpy()
        xs = np.ascontiguousarray(xs[..., 0] + 1j * xs[..., 1]).astype(np.complex64)
        a_gpu = audio[:2].cpu().numpy().astype(np.float64)
        db_gpu = db[:2 * fpb].cpu().numpy().astype(np.float64)
        rms, dberr = 0.0, 0.0
        for b in range(2):
            ref = O.demod(xs[b], FS, args.mode)
            rms = max(rms, float(np.sqrt(np.mean((a_gpu[b] - ref) ** 2))))
            for f in range(fpb):
                want = O.psd_epilogue(O.psd_db(xs[b, f * N_FFT:(f + 1) * N_FFT]))
                dberr = max(dberr, float(np.max(np.abs(db_gpu[b * fpb + f] - want))))
        parity = {"audio_rms_err": rms, "audio_tol": 1e-5, "psd_max_db_err": dberr, "psd_tol_db": 1e-4,
                  "checked_blocks": 2, "ok": bool(rms <= 1e-5 and dberr <= 1e-4)}

    # ---- e2e: host-pointer C ABI, pinned host buffers, copies in the timed region (every rank)
    orig_affinity, numa_note = bind_to_gpu_numa_node(local_rank)
    nb_e = nb
    host_iq = ctx.pinned_empty((nb_e, N_BLOCK), np.complex64)
    host_iq.view(np.float32).reshape(nb_e, N_BLOCK, 2)[:] = iq[:nb_e].cpu().numpy()
    outs = {"audio": ctx.pinned_empty((nb_e, plan.out_len, plan.channels)),
            "cols": ctx.pinned_empty((nb_e * fpb, W_COLS)), "stats": ctx.pinned_empty((nb_e * fpb, 4)),
            "norm": ctx.pinned_empty((nb_e, ROWS_MAX, W_COLS)), "minmax": ctx.pinned_empty((nb_e, 2))}
    ctx.set_stream(None)
    ctx.pipeline(host_iq, FS, args.mode, N_FFT, W_COLS, ROWS_MAX, out=outs)          # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        ctx.pipeline(host_iq, FS, args.mode, N_FFT, W_COLS, ROWS_MAX, out=outs)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / args.e2e_steps * 1e3
    te = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    h2d = host_iq.nbytes
    d2h = sum(v.nbytes for v in outs.values())
    e2e = {"value": world * nb_e * N_BLOCK / e2e_ms / 1e3, "unit": "Msamples/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": args.e2e_steps,
           "api": "pss_pipeline_c64 (host pointers, pinned)", "host_placement": numa_note,
           "audio_matches_device_path": bool(np.array_equal(outs["audio"][:2], audio[:2].cpu().numpy()))}

    if orig_affinity:
        os.sched_setaffinity(0, orig_affinity)               # the CPU baseline uses every host core
    if rank == 0 and not args.no_cpu:
        hb = np.ascontiguousarray(host_iq[:args.cpu_blocks])
        ref = CpuReference(hb, args.mode)
        wall = ref.step()
        ref.close()
        cpu_base = {"value": len(hb) * N_BLOCK / wall / 1e6, "unit": "Msamples/s", "cores": ref.cores, "kind": "port",
                    "sample": f"first {len(hb)} blocks x {N_BLOCK} samples of this run's input, numpy/scipy oracle "
                              f"port, one forked worker per core, wall {wall:.2f} s = {wall * ref.cores:.0f} core-s"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "mode": args.mode, "blocks_per_gpu_per_step": nb,
                       "samples_per_gpu_per_step": nb * N_BLOCK, "fft": N_FFT, "W": W_COLS,
                       "l2_policy": "input per step is 1 GiB per GPU, larger than the 126 MB L2 (no flush needed)",
                       "parallelism": f"frame-sharded x{world}, no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(ln.item()),
            "clocks": clock_info, "parity": parity,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
