#!/usr/bin/env python
"""Summaries of ncu outputs: `launches <csv>` aggregates a gpu__time_duration launch list by kernel;
`raw <ncu-rep>` prints the headline metrics of every captured launch."""
import collections
import csv
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum', 'smsp__inst_executed.sum',
        'smsp__inst_executed_op_shared_atom.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']


def short(name):
    name = name.replace('<unnamed>::', '')
    name = re.sub(r'^void ', '', name)
    return re.sub(r'[<(].*', '', name)


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'us' else v / 1e6 if r[ui] == 'ns' else v
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':40s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:40s} {n:8d} {ms:10.3f} {100 * ms / tot:6.1f}%")
    print(f"{'TOTAL':40s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}")


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    print('kernels:', [short(r[ki]) for r in rows[2:]])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {[r[i] for r in rows[2:]]} {units[i]}")


if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](sys.argv[2])
