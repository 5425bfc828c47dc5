#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into text: tools/ncu_summary.py rep [kernel-regex]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'smsp__inst_executed_op_branch.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'lts__t_bytes.sum']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index('Kernel Name')
    if kre:
        data = [r for r in data if kre in r[kn]]
    print('# kernels:', [r[kn][:60] for r in data])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print('%-75s %-12s %s' % (k, units[i], [r[i] for r in data]))
    print('# warp stall reasons (warps stalled per issue-active cycle), first launch')
    for i, k in enumerate(hdr):
        if 'issue_stalled' in k and 'ratio' in k and 'not_issued' not in k:
            print('%-90s %s' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), data[0][i]))
    src = run(['-i', rep, '--page', 'source', '--csv'] + (['--kernel-name', 'regex:' + kre] if kre else []))
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    if not hi:
        return
    h = rows[hi[0]]
    d = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
    ci = {k: i for i, k in enumerate(h)}

    def f(r, k):
        try:
            return float(r[ci[k]])
        except Exception:
            return 0.0
    tot = sum(f(r, 'Instructions Executed') for r in d)
    ts = sum(f(r, '# Samples') for r in d)
    print('# source page: %d SASS instructions, %.0f warp instructions executed, %d samples' % (len(d), tot, ts))
    regions, cur = [], None
    for idx, r in enumerate(d):
        e = f(r, 'Instructions Executed')
        if cur is None or abs(e - cur[2]) > 0.02 * max(e, cur[2], 1):
            cur = [idx, idx, e, 0.0, 0.0, 0.0]
            regions.append(cur)
        cur[1] = idx
        cur[3] += e
        cur[4] += f(r, 'Thread Instructions Executed')
        cur[5] += f(r, '# Samples')
    print('# regions of equal execution count (>= 0.7 %% of executed instructions)')
    for c in regions:
        if c[3] >= 0.007 * tot:
            print('sass[%5d..%5d] n=%4d exec/instr=%8.0f share=%5.1f%% lanes=%4.1f samples=%4.1f%%  %s' % (
                c[0], c[1], c[1] - c[0] + 1, c[2], 100 * c[3] / tot, c[4] / max(c[3], 1), 100 * c[5] / max(ts, 1),
                d[c[0]][ci['Source']][:50]))
    print('# top 25 SASS instructions by stall samples')
    for r in sorted(d, key=lambda r: -f(r, '# Samples'))[:25]:
        reasons = sorted(((f(r, k), k) for k in ci if k.startswith('stall_') and 'Not Issued' not in k), reverse=True)[:2]
        print('%5d  %-70s %s' % (f(r, '# Samples'), r[ci['Source']][:70], ' '.join('%s=%d' % (k[6:], v) for v, k in reasons if v)))


if __name__ == '__main__':
    main()
