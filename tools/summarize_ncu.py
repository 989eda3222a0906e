#!/usr/bin/env python
"""Summarise .ncu-rep files (read here with `ncu -i`) into small tracked text files under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("gpc__cycles_elapsed.avg.per_second", "sm_clock"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu summary of {rep.split('/')[-1]}\n{note}\n")
        f.write("(ncu --set full --clock-control none; per-launch values; times are cold-cache, serialised)\n\n")
        for r in rows[2:]:
            f.write(f"## {r[idx['Kernel Name']]}\n")
            for k, nm in KEYS:
                if k in idx:
                    f.write(f"{nm:22s} {r[idx[k]]} {units[idx[k]]}\n")
            stalls = []
            for h, i in idx.items():
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") \
                        and "not_issued" not in h:
                    try:
                        stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "")
                                       .replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            f.write("stalls per issue      " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:7]) + "\n\n")


if __name__ == "__main__":
    main()
