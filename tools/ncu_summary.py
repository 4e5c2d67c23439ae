#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the few counters DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [title] > profiles/rNN_ncu_full_x.md
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {title}")
    print(f"# source: ncu -i {rep.split('/')[-1]} --page raw --csv (ncu --set full --clock-control none --import-source on)\n")
    for n, r in enumerate(data):
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"## launch {n}: {d.get('Kernel Name', '')[:140]}")
        for k in KEEP:
            if k in d:
                print(f"{k}: {d[k]} {u[k]}")
        print()


if __name__ == "__main__":
    main()
