#!/usr/bin/env python
"""Digest an .ncu-rep (raw + SASS source pages) into the numbers DESIGN.md / profiles/ quote.

    python tools/ncu_digest.py gpurun_out/prof.ncu-rep [n_tasks] [chunk]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def page(rep, name, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv', *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    n_tasks = int(sys.argv[2]) if len(sys.argv) > 2 else 11264
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    rows = page(rep, 'raw')
    hdr = rows[0]
    for r in rows[2:3]:
        for k in KEYS:
            if k in hdr:
                print(f'{k:70s} {r[hdr.index(k)]} {rows[1][hdr.index(k)]}')
    rows = page(rep, 'source', ['--print-source', 'sass'])
    secs, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'rows': []}
            secs.append(cur)
        elif r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and r:
            cur['rows'].append(r)
    s = secs[0]
    h = s['hdr']
    ie, isrc, ism = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
    stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]

    def op(r):
        t = r[isrc].split()
        return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]

    by, stalls, tot = collections.Counter(), collections.Counter(), 0
    for r in s['rows']:
        by[op(r)] += int(r[ie])
        tot += int(r[ie])
        for i in stall_cols:
            stalls[h[i]] += int(r[i])
    print(f'static SASS instructions {len(s["rows"])}; executed warp-instructions {tot} = {tot / n_tasks:.0f} per task')
    print('per task by opcode:', ', '.join(f'{k} {v / n_tasks:.0f}' for k, v in by.most_common(24)))
    print('stall samples:', ', '.join(f'{k[6:]} {v}' for k, v in stalls.most_common(12)))
    R = s['rows']
    names = ['stall_no_inst', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_mio', 'stall_math', 'stall_not_selected', 'stall_barrier', 'stall_lg', 'stall_dispatch']
    idx = {n: h.index(n) for n in names if n in h}
    for i in range(0, len(R), chunk):
        c = R[i:i + chunk]
        ex = sum(int(r[ie]) for r in c) / n_tasks
        sm = sum(int(r[ism]) for r in c)
        if sm == 0 and ex < 1:
            continue
        ops = collections.Counter(op(r) for r in c)
        st = ' '.join(f'{n[6:]}={sum(int(r[j]) for r in c)}' for n, j in idx.items())
        print(f'{i:5d} exec/task {ex:7.1f} samples {sm:5d} | {st} | ' + ' '.join(f'{k}:{v}' for k, v in ops.most_common(6)))


if __name__ == '__main__':
    main()
