"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: ms, share, launches.
usage: python tools/summarize_launches.py launches.csv > summary.csv"""
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg, total, n = {}, 0.0, 0
for r in rows[1:]:
    name = r[ki]
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'<unnamed>::', '', name)
    name = re.sub(r'\(int\)', '', name)
    name = name.split('(')[0][:80]
    v = float(r[vi].replace(',', ''))
    ms = v / 1e6 if r[ui] == 'ns' else (v / 1e3 if r[ui] == 'us' else v)
    a = agg.setdefault(name, [0.0, 0])
    a[0] += ms; a[1] += 1
    total += ms; n += 1
ours = {k: v for k, v in agg.items() if not k.startswith(('at::', 'magma', 'cutlass', 'ampere', 'sm', 'nccl', 'cublas', 'void at', 'gemv', 'std::'))}
print(f'# total {total:.2f} ms in {n} launches; libfreqair kernels {sum(v[0] for v in ours.values()):.2f} ms '
      f'({100 * sum(v[0] for v in ours.values()) / total:.1f} %) in {sum(v[1] for v in ours.values())} launches')
print('ms,share_pct,launches,kernel')
for k, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{ms:.3f},{100 * ms / total:.2f},{c},"{k}"')
