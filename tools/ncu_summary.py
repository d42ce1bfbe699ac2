#!/usr/bin/env python
"""Prints the metrics we track from an .ncu-rep (first kernel in the report): python tools/ncu_summary.py file.ncu-rep [json-out]"""
import csv, io, json, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_op_ldgsts.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__grid_size', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_imc_miss_per_warp_active.pct', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
        'smsp__warp_issue_stalled_membar_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__inst_executed_pipe_uniform.sum', 'idc__requests.sum', 'idc__requests_lookup_miss.sum', 'sm__inst_executed_pipe_adu.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
res = {'kernel': vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else None}
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        res[w] = [vals[i], units[i]]
        print(f"{w:90s} {vals[i]:>18s} {units[i]}")
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], 'w'), indent=1)
