#!/usr/bin/env python
"""Text summary of an .ncu-rep (one column per captured launch) with the metrics the roofline / DESIGN.md tables quote.
Usage: ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__m_l1tex2xbar_write_bytes.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
print("Kernel Name []:", " | ".join(r[ix["Kernel Name"]][:60] for r in data))
for k in KEYS:
    if k in ix:
        print(f"{k} [{units[ix[k]]}]:", " | ".join(r[ix[k]] for r in data))
