#!/usr/bin/env python
"""Condensed view of an .ncu-rep (read here, no GPU): headline metrics, opcode mix, hottest source lines."""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum ', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active', 'smsp__average_warps_issue_stalled_wait_per_issue_active',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active',
        'smsp__average_warps_issue_stalled_membar_per_issue_active', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active',
        'lts__t_sectors_op_read.sum ', 'lts__t_sectors_op_write.sum ', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum ']


def I(x):
    """ncu leaves cells empty for lines without samples and writes thousands separators in some locales"""
    try:
        return int(float(str(x).replace(',', '') or 0))
    except ValueError:
        return 0


def page(rep, name):
    return subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout


def main(rep, nlines=25):
    raw = list(csv.reader(io.StringIO(page(rep, 'raw'))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    print('kernel:', vals[hdr.index('Kernel Name')][:150])
    for h, u, v in zip(hdr, units, vals):
        if any((h + ' ').startswith(k) or h == k.strip() for k in KEYS):
            print('  %-86s %-10s %s' % (h[:86], u, v))
    src = list(csv.reader(io.StringIO(page(rep, 'source'))))
    h2 = src[1]
    isrc, ins, isamp = h2.index('Source'), h2.index('Instructions Executed'), h2.index('# Samples')
    data = src[2:]
    data = [r for r in data if len(r) > max(isrc, ins, isamp)]
    tot = sum(I(r[ins]) for r in data)
    tsamp = sum(I(r[isamp]) for r in data)
    op = collections.Counter()
    for r in data:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
        op[m.group(2).split('.')[0] if m else '?'] += I(r[ins])
    print('warp-instructions executed: %d ; opcode mix:' % tot, ', '.join('%s %.1f%%' % (k, 100 * v / max(tot, 1)) for k, v in op.most_common(14)))
    print('hottest SASS by stall samples (of %d):' % tsamp)
    stall_cols = [i for i, n in enumerate(h2) if n.startswith('stall_') and 'Not Issued' not in n]
    for r in sorted(data, key=lambda r: -I(r[isamp]))[:nlines]:
        st = sorted(((I(r[i]), h2[i]) for i in stall_cols if r[i] not in ('', '0')), reverse=True)[:2]
        print('  %5.1f%%  exec %9s  %-70s %s' % (100 * I(r[isamp]) / max(tsamp, 1), r[ins], r[isrc].strip()[:70], st))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
