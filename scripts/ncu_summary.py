#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics per kernel and, with --sass, the
per-instruction execution counts of one kernel grouped by how often each instruction ran.
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--sass REGEX] [--per N]"""
import argparse, csv, io, subprocess, sys, collections

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep')
    ap.add_argument('--sass', default=None)
    ap.add_argument('--per', type=float, default=1.0, help='divide instruction counts by this (e.g. rows)')
    ap.add_argument('--full', action='store_true')
    a = ap.parse_args()
    rows = list(csv.reader(io.StringIO(ncu(['-i', a.rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('----', d['Kernel Name'][:100])
        for w in WANT:
            if w in d:
                print(f'  {w:86s} {d[w]:>16s} {units[hdr.index(w)]}')
    if a.sass:
        rows = list(csv.reader(io.StringIO(ncu(['-i', a.rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + a.sass]))))
        hdr = rows[1]
        ia, isrc, isamp, ith = (hdr.index(x) for x in ('Instructions Executed', 'Source', '# Samples', 'Avg. Threads Executed'))
        data = rows[2:]
        tot = sum(int(r[ia]) for r in data)
        tots = max(1, sum(int(r[isamp]) for r in data))
        print(f'total warp instructions {tot}  ({tot / a.per:.1f} per unit), {len(data)} SASS instructions')
        grp = collections.OrderedDict()
        for r in data:
            k = round(int(r[ia]) / a.per, 1)
            g = grp.setdefault(k, [0, 0])
            g[0] += 1
            g[1] += int(r[isamp])
        print('  exec/unit  #sass  warp-inst/unit  samples%')
        for k, g in sorted(grp.items()):
            if k * g[0] >= 0.005 * tot / a.per:
                print(f'  {k:9.1f} {g[0]:6d} {k * g[0]:12.1f} {100 * g[1] / tots:8.1f}')
        if a.full:
            for i, r in enumerate(data):
                print(f'{i:4d} {int(r[ia]) / a.per:9.1f} {100 * int(r[isamp]) / tots:5.1f}% th={r[ith]:>5s} {r[isrc].strip()[:100]}')


if __name__ == '__main__':
    main()
