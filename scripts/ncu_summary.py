"""Summarise an ncu report (raw page CSV on stdin) into the metrics DESIGN.md /
profiles/ quote.  Usage:
    ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_summary.py
"""
import csv
import sys

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum',
    'lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.sum',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__grid_size',
    'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__maximum_warps_per_active_cycle_pct',
    'launch__waves_per_multiprocessor', 'smsp__cycles_active.avg',
    'launch__shared_mem_per_block_static', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'smsp__inst_executed_op_shared_ld.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum',
    'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fmaheavy.sum',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_cbu.sum',
]

rows = list(csv.reader(sys.stdin))
hdr = rows[0]
units = rows[1]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print('== kernel:', name)
    for i, h in enumerate(hdr):
        if h in WANT:
            print('  %-70s %-12s %s' % (h, units[i], r[i]))
    stalls = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
              if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and r[i]]
    if not stalls:
        stalls = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
                  if 'warp_issue_stalled' in h and h.endswith('.pct') and r[i]]
    for v, h in sorted(stalls, reverse=True)[:8]:
        print('  stall %-64s %.3f' % (h, v))
