#!/usr/bin/env python
"""Copy what the judge should read from gpurun_out/ (scratch) into profiles/ (tracked):
the launch list CSV as is, and for every .ncu-rep a short text summary of the raw page.

    python tools/summarise_profiles.py <tag>       # e.g. r1a
"""
import csv
import glob
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main():
    tag = sys.argv[1]
    src = os.path.join(ROOT, "gpurun_out")
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    for f in glob.glob(os.path.join(src, "launches_%s.csv" % tag)):
        lines = [l for l in open(f) if l.startswith('"')]
        open(os.path.join(dst, os.path.basename(f)), "w").writelines(lines)
    for rep in sorted(glob.glob(os.path.join(src, "prof_*_%s.ncu-rep" % tag))):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        rows = [r for r in rows if len(r) > 10]
        hdr, units = rows[0], rows[1]
        name = os.path.basename(rep).replace(".ncu-rep", ".txt")
        with open(os.path.join(dst, name), "w") as fo:
            fo.write("# ncu --set full --clock-control none, raw page excerpt of %s\n" % os.path.basename(rep))
            for vals in rows[2:]:
                kn = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
                fo.write("kernel: %s\n" % kn)
                for w in WANT:
                    if w in hdr:
                        i = hdr.index(w)
                        fo.write("  %-90s %s %s\n" % (w, vals[i], units[i]))
        print("wrote", name)


if __name__ == "__main__":
    main()
