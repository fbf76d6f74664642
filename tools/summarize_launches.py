"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel: ms, share, launches and (when present) DRAM bytes per launch.
usage: python tools/summarize_launches.py launches.csv [traffic.json] > summary.csv
With a second argument the DRAM traffic of the contraction class and of the depthwise-conv + LayerNorm class is written
as JSON (bench.py copies it into roofline.traffic / roofline_hbm.traffic)."""
import csv
import json
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui, mi, idi = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Metric Name', 'ID'))
launch = {}
for r in rows[1:]:
    name = re.sub(r'\(int\)', '', re.sub(r'<unnamed>::', '', re.sub(r'^void ', '', r[ki]))).split('(')[0][:80]
    v = float(r[vi].replace(',', ''))
    d = launch.setdefault(r[idi], {'name': name, 'ms': 0.0, 'bytes': 0.0})
    if r[mi].startswith('gpu__time_duration'):
        d['ms'] = v / 1e6 if r[ui] in ('ns', 'nsecond') else (v / 1e3 if r[ui] in ('us', 'usecond') else v)
    elif r[mi].startswith('dram__bytes'):
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[ui], 1)
        d['bytes'] += v * scale
agg, total = {}, 0.0
for d in launch.values():
    a = agg.setdefault(d['name'], [0.0, 0, 0.0])
    a[0] += d['ms']; a[1] += 1; a[2] += d['bytes']
    total += d['ms']
n = len(launch)
lib = ('at::', 'magma', 'cutlass', 'ampere', 'sm', 'nccl', 'cublas', 'void at', 'gemv', 'std::')
ours = {k: v for k, v in agg.items() if not k.startswith(lib)}
print(f'# total {total:.2f} ms in {n} launches; libfreqair kernels {sum(v[0] for v in ours.values()):.2f} ms '
      f'({100 * sum(v[0] for v in ours.values()) / total:.1f} %) in {sum(v[1] for v in ours.values())} launches')
print('ms,share_pct,launches,dram_MB_per_launch,kernel')
for k, (ms, c, by) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{ms:.3f},{100 * ms / total:.2f},{c},{by / c / 1e6:.2f},"{k}"')
if len(sys.argv) > 2:
    def cls(pred):
        sel = [v for k, v in agg.items() if pred(k)]
        c = sum(v[1] for v in sel)
        return {'launches': c, 'dram_bytes_per_launch': sum(v[2] for v in sel) / max(c, 1), 'ms': sum(v[0] for v in sel)}
    json.dump({'gemm': cls(lambda k: k.startswith(('gemm_tc_kernel', 'gemm_simt'))),
               'dwconv_ln': cls(lambda k: k.startswith(('dwconv', 'layernorm', 'ln_'))),
               'note': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over every launch of the class in ONE '
                       'train step (ncu, serialised launches, cold L2 between kernels): tools/one_step.py'}, open(sys.argv[2], 'w'), indent=1)
