#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key per-launch metrics as JSON lines."""
import csv, subprocess, sys, json
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread',
 'launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed.sum','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__cycles_active.avg','sm__cycles_elapsed.avg',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct',
 'smsp__warp_issue_stalled_drain_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct']
out = subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = {'kernel': r[hdr.index('Kernel Name')][:50]}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w); d[w] = r[i] + ' ' + units[i]
    print(json.dumps(d, indent=1))
