#!/usr/bin/env python
"""Summarise an ncu report: per-kernel headline metrics + barrier positions + hottest SASS lines.
usage: python tools/ncu_hot.py report.ncu-rep [kernel-substring] [top-n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 else ''; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 14
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor']
for r in rows[2:]:
    if filt in r[hdr.index('Kernel Name')]:
        print({w.split('.')[0][-28:] + '.' + w.split('.')[-1][:12]: r[hdr.index(w)][:48] for w in want if w in hdr})
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern = None; data = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'Kernel Name':
        kern = r[1] + '#' + str(len(data)); data[kern] = []; continue
    if kern: data[kern].append(r)
seen = set()
for k, v in data.items():
    base = k.split('#')[0]
    if filt not in k or base in seen: continue
    seen.add(base)
    hdr = v[0]; body = v[1:]
    isamp = hdr.index('# Samples'); isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed')
    tot = sum(int(r[isamp]) for r in body) or 1
    print('\n==', base[:100], 'instrs', len(body), 'samples', tot)
    cum = 0
    for i, r in enumerate(body):
        cum += int(r[isamp])
        if 'BAR' in r[isrc]: print('   bar @%d cum %.1f%% exec %s' % (i, 100 * cum / tot, r[iex]))
    for i, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][isamp]))[:topn]:
        print('     %5d %6s %8s  %s' % (i, r[isamp], r[iex], r[isrc].strip()[:90]))
    cat = {}
    for r in body:
        t = r[isrc].strip().split()
        op = t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '')
        cat[op] = cat.get(op, 0) + int(r[isamp])
    print('   by opcode:', sorted(cat.items(), key=lambda kv: -kv[1])[:10])
