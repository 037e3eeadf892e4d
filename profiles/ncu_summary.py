#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md argues from.
Usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [kernel-substring] > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_fp64.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum']
STALLS = 'smsp__average_warps_issue_stalled_'


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ''
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index('Kernel Name')
    for r in rows[2:]:
        if sub and sub not in r[name_col]:
            continue
        print('kernel:', r[name_col])
        d = dict(zip(hdr, zip(units, r)))
        for k in KEEP:
            if k in d:
                print('  %-75s %s %s' % (k, d[k][1], d[k][0]))
        st = sorted(((float(v[1] or 0), k) for k, v in d.items()
                     if k.startswith(STALLS) and k.endswith('_per_issue_active.ratio')), reverse=True)
        print('  stall reasons (warps stalled per issued instruction):')
        for v, k in st[:8]:
            print('    %-40s %.3f' % (k[len(STALLS):-len('_per_issue_active.ratio')], v))


if __name__ == '__main__':
    main()
