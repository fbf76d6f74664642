#!/usr/bin/env python
"""Per-kernel-call time attribution of one configs[1] train step (CUDA events around every C-ABI call).
Serialises nothing (events on the launching stream), so shares are faithful; absolute times include launch gaps when
the GPU runs ahead of the host.  Usage (GPU box): python tools/profile_step.py [--batch 16] [--top 60]"""
import argparse
import collections
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

PKG = bench.PKG


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--top', type=int, default=60)
    ap.add_argument('--gemm', action='store_true')
    ap.add_argument('--workload', default='train', choices=['train', 'vit_dgrn_train'])
    a = ap.parse_args()
    ops = importlib.import_module(PKG + '.ops')
    synth = importlib.import_module(PKG + '.synth')
    model = importlib.import_module(PKG + '.net.model')
    trainer = importlib.import_module(PKG + '.trainer')
    torch.manual_seed(0)
    net = model.AirNet(bench.make_opt(a.batch, a.workload)).cuda().train()
    ts = trainer.TrainStep(net, lr=2e-4, contrast_loss_weight=0.6)
    x = [t.cuda() for t in (synth.mixed_batch(a.batch) if a.workload == 'vit_dgrn_train' else synth.noisy_batch(a.batch, 25))]
    for _ in range(3):
        ts.step(*x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ts.step(*x); e1.record(); torch.cuda.synchronize()
    print(f'step (unprofiled): {e0.elapsed_time(e1):.2f} ms')
    ops.PROFILE = []
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(); ts.step(*x); f1.record(); torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    print(f'step (profiled): {f0.elapsed_time(f1):.2f} ms, {len(rec)} library calls')
    by_name = collections.defaultdict(lambda: [0.0, 0])
    by_sig = collections.defaultdict(lambda: [0.0, 0])
    for name, sig, a0, a1 in rec:
        ms = a0.elapsed_time(a1)
        by_name[name][0] += ms; by_name[name][1] += 1
        key = (name, sig[:8] if name == 'fa_gemm' else sig[:6])
        by_sig[key][0] += ms; by_sig[key][1] += 1
    tot = sum(v[0] for v in by_name.values())
    print(f'sum of library-call times: {tot:.2f} ms')
    for n, (ms, c) in sorted(by_name.items(), key=lambda kv: -kv[1][0]):
        print(f'{ms:9.3f} ms {100 * ms / tot:5.1f}% {c:5d}  {n}')
    if a.gemm:
        print('--- every fa_gemm signature: ms, calls, TFLOP/s, GB/s (M N K lda ldb ldc tA tB)')
        for (n, sig), (ms, c) in sorted(by_sig.items(), key=lambda kv: -kv[1][0]):
            if n != 'fa_gemm':
                continue
            M, N, K = sig[0], sig[1], sig[2]
            fl = 2.0 * M * N * K * c
            by = 4.0 * (M * K + N * K + M * N) * c
            print(f'{ms:9.3f} ms {c:4d}  {fl / ms / 1e9:7.1f} TF/s {by / ms / 1e6:7.0f} GB/s  {sig}')
    print('--- by signature (gemm: M N K lda ldb ldc tA tB)')
    for (n, sig), (ms, c) in sorted(by_sig.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f'{ms:9.3f} ms {100 * ms / tot:5.1f}% {c:4d}  {n} {sig}')


if __name__ == '__main__':
    main()
