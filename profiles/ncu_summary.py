#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the handful of metrics the roofline
discussion needs.   python profiles/ncu_summary.py gpurun_out/x.ncu-rep [substring filters...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_fma", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_cycles_active", "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__warp_issue_stalled", "smsp__pcsamp_warps_issue_stalled"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"=== {name[:100]}  block {r[hdr.index('Block Size')]} grid {r[hdr.index('Grid Size')]}")
        for i, h in enumerate(hdr):
            short = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
            if any(k in h for k in KEYS + extra):
                print(f"  {h:100s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
