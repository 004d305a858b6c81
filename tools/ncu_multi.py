#!/usr/bin/env python
"""Side-by-side headline metrics of every kernel in an .ncu-rep (read here, no GPU) + hottest SASS lines per kernel.
usage: ncu_multi.py report.ncu-rep [n_hot_lines]"""
import csv
import io
import subprocess
import sys

KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_active.avg',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size']
STALLS = ['long_scoreboard', 'short_scoreboard', 'mio_throttle', 'barrier', 'wait', 'math_pipe_throttle', 'lg_throttle',
          'not_selected', 'dispatch_stall', 'no_instruction', 'branch_resolving', 'membar', 'sleeping', 'tex_throttle',
          'drain', 'imc_miss', 'selected']


def page(rep, name, extra=()):
    return subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'] + list(extra), capture_output=True, text=True).stdout


def main(rep, nhot=0):
    raw = list(csv.reader(io.StringIO(page(rep, 'raw'))))
    hdr = raw[0]
    keys = KEYS + ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s for s in STALLS]
    for k in keys:
        idx = [i for i, h in enumerate(hdr) if h == k]
        if not idx:
            continue
        i = idx[0]
        print('%-78s %s  [%s]' % (k.replace('smsp__average_warps_issue_stalled_', 'stall:')[:78],
                                  ' | '.join('%14s' % r[i][:14] for r in raw[2:]), raw[1][i]))
    if nhot:
        for kid in range(len(raw) - 2):
            src = list(csv.reader(io.StringIO(page(rep, 'source', ['--print-source', 'sass', '--kernel-id', '::%d' % kid]))))
            print(src[:3])


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
