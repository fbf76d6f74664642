#!/usr/bin/env python
"""Micro-benchmark of the library's kernels on the shapes of the configs[1] train step (B=16, 128x128 crops).
Prints one line per case: CUDA-event time (median of `reps`, L2 flushed between launches), TFLOP/s and GB/s of the
algorithmic bytes.  Usage (GPU box):  python tools/bench_kernels.py gemm [--backend 0|1|2]
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = 'frequency-wised_all-in-one_image_restoration_model_b200'
ops = importlib.import_module(PKG + '.ops')

FLUSH = None


def timeit(fn, reps=7, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        if flush:
            FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def gemm_cases():
    B = 16
    cases = []
    # decoder levels: (res, dim)
    for res, dim in [(128, 56), (64, 112), (32, 224), (16, 448), (8, 896), (16, 896), (32, 448), (64, 224), (128, 112)]:
        T = B * res * res
        cases += [(f'dec{res}/{dim} q', T, dim, dim), (f'dec{res}/{dim} kv', T, 2 * dim, dim),
                  (f'dec{res}/{dim} leff1', T, 4 * dim, dim), (f'dec{res}/{dim} leff2', T, dim, 4 * dim)]
    for res, dim in [(128, 28), (64, 56), (32, 112), (16, 224), (8, 448)]:
        T = 3 * B * res * res
        cases += [(f'enc{res}/{dim} q', T, dim, dim), (f'enc{res}/{dim} kv', T, 2 * dim, dim), (f'enc{res}/{dim} leff1', T, 4 * dim, dim)]
    cases += [('head 448->65536', B * 64, 65536, 448)]
    return cases


def bench_gemm(backend, only=None, layouts='NT,NN,TN', bexact=False):
    print(f'{"case":24s} {"layout":3s} {"M":>8s} {"N":>6s} {"K":>6s} {"ms":>8s} {"TFLOP/s":>8s} {"GB/s":>8s}')
    for name, M, N, K in gemm_cases():
        if only and only not in name:
            continue
        X = torch.randn(M, K, device='cuda')
        W = torch.randn(N, K, device='cuda') * 0.05
        Y = torch.empty(M, N, device='cuda')
        dX = torch.empty(M, K, device='cuda')
        dW = torch.zeros(N, K, device='cuda')
        for lay, fn, flops, byts in [
            ('NT', lambda: ops.gemm(X, W, Y, backend=backend, b_is_tf32=bexact), 2 * M * N * K, 4 * (M * K + N * K + M * N)),
            ('NN', lambda: ops.gemm(Y, W, dX, transB=False, backend=backend, b_is_tf32=bexact), 2 * M * N * K, 4 * (M * N + N * K + M * K)),
            ('TN', lambda: ops.gemm(Y, X, dW, transA=True, transB=False, accumulate=True, backend=backend), 2 * M * N * K,
             4 * (M * N + M * K + N * K)),
        ]:
            if lay not in layouts.split(','):
                continue
            try:
                ms = timeit(fn)
            except RuntimeError as e:
                print(f'{name:24s} {lay:3s} {M:8d} {N:6d} {K:6d}   error: {str(e)[:80]}')
                continue
            print(f'{name:24s} {lay:3s} {M:8d} {N:6d} {K:6d} {ms:8.3f} {flops / ms / 1e9:8.1f} {byts / ms / 1e6:8.0f}', flush=True)
        del X, W, Y, dX, dW


def bench_epi(backend):
    """Epilogue flavours on the widest LeFF shape of the decoder (T = 262144 tokens, C = 112, hidden 448)."""
    M, C, Hd = 262144, 112, 448
    X = torch.randn(M, C, device='cuda'); W1 = torch.randn(Hd, C, device='cuda') * 0.05; b1 = torch.randn(Hd, device='cuda')
    U = torch.empty(M, Hd, device='cuda'); Hh = torch.empty(M, Hd, device='cuda')
    W2 = torch.randn(C, Hd, device='cuda') * 0.05; b2 = torch.randn(C, device='cuda')
    Y = torch.empty(M, C, device='cuda'); R = torch.randn(M, C, device='cuda'); rs = torch.ones(16, device='cuda')
    G = torch.randn(M, C, device='cuda'); dU = torch.empty(M, Hd, device='cuda')
    cases = [
        ('leff1 plain', lambda: ops.gemm(X, W1, Hh, backend=backend), 4 * (M * C + M * Hd)),
        ('leff1 +bias', lambda: ops.gemm(X, W1, Hh, bias=b1, backend=backend), 4 * (M * C + M * Hd)),
        ('leff1 +bias+gelu', lambda: ops.gemm(X, W1, Hh, bias=b1, act=ops.ACT_GELU, backend=backend), 4 * (M * C + M * Hd)),
        ('leff1 +bias+gelu+preact', lambda: ops.gemm(X, W1, Hh, bias=b1, act=ops.ACT_GELU, preact=U, backend=backend), 4 * (M * C + 2 * M * Hd)),
        ('leff2 plain', lambda: ops.gemm(Hh, W2, Y, backend=backend), 4 * (M * C + M * Hd)),
        ('leff2 +bias+rs+res', lambda: ops.gemm(Hh, W2, Y, bias=b2, rowscale=rs, rows_per_scale=16384, residual=R, backend=backend), 4 * (2 * M * C + M * Hd)),
        ('dX2 plain (NN)', lambda: ops.gemm(G, W2, dU, transB=False, backend=backend), 4 * (M * C + M * Hd)),
        ('dX2 *gelu\'(aux) (NN)', lambda: ops.gemm(G, W2, dU, transB=False, aux=U, aux_act=ops.ACT_GELU, backend=backend), 4 * (M * C + 2 * M * Hd)),
        ('dX2 *aux (NN, ACT_MUL)', lambda: ops.gemm(G, W2, dU, transB=False, aux=U, aux_act=ops.ACT_MUL, backend=backend), 4 * (M * C + 2 * M * Hd)),
        ('dX1 accumulate (NN)', lambda: ops.gemm(dU, W1, Y, transB=False, accumulate=True, backend=backend), 4 * (2 * M * C + M * Hd)),
    ]
    for name, fn, byts in cases:
        ms = timeit(fn)
        print(f'{name:28s} {ms:8.3f} ms {byts / ms / 1e6:8.0f} GB/s', flush=True)


def bench_misc():
    """Bandwidth kernels at the shapes of the step: algorithmic GB/s against the 6.5 TB/s HBM roof."""
    print(f'{"kernel":44s} {"ms":>8s} {"GB/s":>8s}')

    def rep(name, fn, byts):
        ms = timeit(fn)
        print(f'{name:44s} {ms:8.3f} {byts / ms / 1e6:8.0f}', flush=True)
    for rows, C in [(262144, 56), (262144, 112), (786432, 28), (65536, 224), (16384, 448), (4096, 896), (1024, 896)]:
        x = torch.randn(rows, C, device='cuda'); g = torch.randn(C, device='cuda'); b = torch.randn(C, device='cuda')
        y = torch.empty_like(x); dy = torch.randn_like(x); dres = torch.randn_like(x); dx = torch.empty_like(x)
        dg = torch.zeros(C, device='cuda'); db = torch.zeros(C, device='cuda')
        _, mean, rstd = ops.layernorm_fwd(x, g, b, y)
        rep(f'layernorm_fwd {rows}x{C}', lambda: ops.layernorm_fwd(x, g, b, y), 8 * rows * C)
        rep(f'layernorm_bwd {rows}x{C}', lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dres, dg, db, dx), 16 * rows * C)
        out = torch.zeros(C, device='cuda')
        rep(f'colsum {rows}x{C}', lambda: ops.colsum(x, out), 4 * rows * C)
        rs = torch.ones(16, device='cuda')
        rep(f'scale_rows {rows}x{C}', lambda: ops.scale_rows(x, rs, rows // 16), 8 * rows * C)
        del x, y, dy, dres, dx
    for B, H, C in [(16, 128, 224), (16, 128, 448), (48, 128, 112), (16, 64, 896), (16, 32, 1792), (16, 16, 3584), (16, 8, 3584)]:
        h1 = torch.randn(B * H * H, C, device='cuda'); w = torch.randn(C, 1, 3, 3, device='cuda'); bb = torch.randn(C, device='cuda')
        n = B * H * H * C
        rep(f'dwconv_fwd B{B} {H}x{H} C{C}', lambda: ops.dwconv_fwd(h1, w, bb, B, H, H, C), 12 * n)
        du2 = torch.randn_like(h1); u1 = torch.randn_like(h1)
        dw = torch.zeros_like(w); dbb = torch.zeros(C, device='cuda')
        rep(f'dwconv_bwd B{B} {H}x{H} C{C}', lambda: ops.dwconv_bwd(du2, h1, u1, w, dw, dbb, B, H, H, C), 16 * n)
        del h1, du2, u1


def bench_attn():
    """Fused window attention kernels at the step's shapes: time, effective TFLOP/s of the contractions (dense flops of
    the unmasked blocks, fwd 2 products, bwd 5 + the recomputed score product) and GB/s of q,k,v,o(,grads)."""
    print(f'{"kernel":46s} {"ms":>8s} {"TFLOP/s":>8s} {"GB/s":>8s}')
    bob = None
    from importlib import import_module
    fd = import_module(PKG + '.net.utils.frequency_decompose')
    for B, H, heads, hd, shift in [(16, 128, 2, 56, 4), (16, 128, 1, 56, 0), (16, 64, 4, 56, 4), (16, 32, 8, 56, 4), (16, 16, 16, 56, 4)]:
        C = heads * hd; T = B * H * H
        qkv = torch.randn(T, 3 * C, device='cuda') * 0.5
        table = torch.randn(225, heads, device='cuda') * 0.3
        coef = torch.randn(B, heads, 3, device='cuda') * 0.3
        if bob is None:
            bob = fd.half_band_map('frequency_decompose_1', 0.5, 64).cuda()
        o = torch.empty(T, C, device='cuda'); dO = torch.randn(T, C, device='cuda')
        dq = torch.empty(T, C, device='cuda'); dkv = torch.empty(T, 2 * C, device='cuda')
        dtab = torch.zeros_like(table); dcf = torch.zeros_like(coef)
        items = B * (H // 8) ** 2 * heads
        fl = items * 2 * 64 * 64 * hd * 2
        args = (B, H, H, heads, hd, shift, hd ** -0.5)
        ms = timeit(lambda: ops.win_attn_fwd(qkv[:, :C], qkv[:, C:], o, *args, table, coef, heads, bob, 3))
        print(f'{f"win_attn_fwd B{B} {H}x{H} h{heads} hd{hd} s{shift}":46s} {ms:8.3f} {fl / ms / 1e9:8.1f} {16 * T * C / ms / 1e6:8.0f}', flush=True)
        ms = timeit(lambda: ops.win_attn_bwd(qkv[:, :C], qkv[:, C:], dO, dq, dkv, *args, table, dtab, coef, heads, dcf, bob, 3))
        print(f'{f"win_attn_bwd B{B} {H}x{H} h{heads} hd{hd} s{shift}":46s} {ms:8.3f} {3 * fl / ms / 1e9:8.1f} {28 * T * C / ms / 1e6:8.0f}', flush=True)
    L = 3
    for B, H, heads in [(16, 128, 1), (16, 64, 2), (16, 32, 4), (16, 16, 8)]:
        hd = 28; C = heads * hd; T = L * B * H * H
        qkv = torch.randn(T, 3 * C, device='cuda') * 0.5
        tables = torch.randn(L * L, 225, heads, device='cuda') * 0.3
        o = torch.empty(T, C, device='cuda'); dO = torch.randn(T, C, device='cuda')
        dq = torch.empty(T, C, device='cuda'); dkv = torch.empty(T, 2 * C, device='cuda'); dtab = torch.zeros_like(tables)
        items = L * B * (H // 8) ** 2 * heads
        for kind, nk in ((0, 64), (1, 128)):
            fl = items * 2 * 64 * nk * hd * 2
            ms = timeit(lambda: ops.joint_attn_fwd(qkv[:, :C], qkv[:, C:], o, L, B, H, H, heads, hd, 4, hd ** -0.5, tables, kind))
            print(f'{f"joint_fwd kind{kind} B{B} {H}x{H} h{heads}":46s} {ms:8.3f} {fl / ms / 1e9:8.1f} {16 * T * C / ms / 1e6:8.0f}', flush=True)
            ms = timeit(lambda: ops.joint_attn_bwd(qkv[:, :C], qkv[:, C:], dO, dq, dkv, L, B, H, H, heads, hd, 4, hd ** -0.5, tables, dtab, kind))
            print(f'{f"joint_bwd kind{kind} B{B} {H}x{H} h{heads}":46s} {ms:8.3f} {3 * fl / ms / 1e9:8.1f} {28 * T * C / ms / 1e6:8.0f}', flush=True)


def bench_dgrn():
    """The DGRN-side kernels at configs[0] / configs[2] shapes (128 x 128, 64 channels) and the standalone band filter (K1),
    BatchNorm (K9) and SFT (K8) kernels: algorithmic GB/s against the HBM roof; the implicit-GEMM convs also in TFLOP/s."""
    from importlib import import_module
    fd = import_module(PKG + '.net.utils.frequency_decompose')
    print(f'{"kernel":52s} {"ms":>8s} {"GB/s":>8s} {"TFLOP/s":>8s}')

    def rep(name, fn, byts, flops=0):
        ms = timeit(fn)
        print(f'{name:52s} {ms:8.3f} {byts / ms / 1e6:8.0f} {flops / ms / 1e9:8.1f}', flush=True)
    for B in (4, 16):
        H = W = 128; C = 64; T = B * H * W
        x = torch.randn(T, C, device='cuda'); om = torch.randn(T, 32, device='cuda') * 0.5
        col = torch.empty(T, 9 * C, device='cuda'); dcol = torch.randn(T, 9 * C, device='cuda')
        rep(f'dcn_im2col B{B} 128x128 C64', lambda: ops.dcn_im2col(x, om, B, H, W, C, col), 4 * T * (C + 32 + 9 * C))
        rep(f'dcn_col2im B{B} 128x128 C64', lambda: ops.dcn_col2im(x, om, dcol, B, H, W, C), 4 * T * (C + 32 + 9 * C + C + 32))
        wk = torch.randn(C, 9 * C, device='cuda') * 0.05; y = torch.empty(T, C, device='cuda'); b = torch.randn(C, device='cuda')
        fl = 2 * T * C * 9 * C
        rep(f'conv3x3 implicit fwd B{B} 64->64 +bias+lrelu', lambda: ops.conv3x3_gemm(x.view(B, H * W, C), wk, y, B, H, W, bias=b, act=ops.ACT_LRELU, act_param=0.1), 8 * T * C, fl)
        dwk = torch.zeros(C, 9 * C, device='cuda')
        rep(f'conv3x3 implicit wgrad B{B} 64->64', lambda: ops.conv3x3_wgrad(y, x.view(B, H * W, C), dwk, B, H, W), 8 * T * C, fl)
        colx = torch.empty(T, 9 * C, device='cuda')
        rep(f'  (explicit: im2col + gemm B{B})', lambda: (ops.im2col(x.view(B, H * W, C), B, H, W, C, 3, 3, 1, 1), ops.gemm(colx, wk, y, bias=b)), 8 * T * C, fl)
        g = torch.randn(T, C, device='cuda'); be = torch.randn(T, C, device='cuda'); d = torch.randn(T, C, device='cuda')
        rep(f'sft_fuse_fwd B{B}', lambda: ops.sft_fuse_fwd(x, d, g, be, 0.1), 20 * T * C)
        rep(f'sft_fuse_bwd B{B}', lambda: ops.sft_fuse_bwd(x, d, g, be, y, 0.1), 36 * T * C)
        del x, om, col, dcol, colx
    # K1 standalone: band split / filter of 64 x 64 attention maps and 128 x 128 images
    for nmaps, n in ((74752 // 4, 64), (48, 128), (4096, 128)):
        xm = torch.rand(nmaps, n, n, device='cuda')
        bob = fd.half_band_map('frequency_decompose_1', 0.5, n).cuda()
        coef = torch.randn(1, 3, device='cuda') * 0.3
        rep(f'band_filter {nmaps} maps {n}x{n}', lambda: ops.band_filter(xm, bob, coef, nmaps, 1), 8 * nmaps * n * n)
        rep(f'band_split (3 bands) {nmaps} maps {n}x{n}', lambda: ops.band_split(xm, bob, 3), 16 * nmaps * n * n)
    # K9: BatchNorm statistics / apply on the encoder-head layout [B, 256, 128*128]
    Bq, Cq, S = 16, 256, 16384
    h = torch.randn(Bq, Cq, S, device='cuda'); sc = torch.rand(Cq, device='cuda'); sh = torch.randn(Cq, device='cuda')
    rep('bn_stats [16,256,16384]', lambda: ops.bn_stats(h, Bq, Cq, S), 4 * h.numel())
    rep('bn_apply (+lrelu+pool) [16,256,16384]', lambda: ops.bn_apply(h, sc, sh, 0.1, Bq, Cq, S), 4 * h.numel())
    pm = torch.randn(99_715_312, device='cuda'); pk = torch.randn_like(pm)
    rep('momentum_update 99.7M params', lambda: ops.momentum_update(pk, pm, 0.999), 12 * pm.numel())
    gr = torch.randn_like(pm); m1 = torch.zeros_like(pm); v1 = torch.zeros_like(pm)
    rep('adam_step 99.7M params', lambda: ops.adam_step(pm, gr, m1, v1, 2e-4, 0.9, 0.999, 1e-8, 3), 28 * pm.numel())
    rep('round_tf32 99.7M params', lambda: ops.round_tf32(pm, pk), 8 * pm.numel())


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', choices=['gemm', 'epi', 'misc', 'attn', 'dgrn'])
    ap.add_argument('--backend', type=int, default=0)
    ap.add_argument('--only', default=None)
    ap.add_argument('--layouts', default='NT,NN,TN')
    ap.add_argument('--bexact', action='store_true', help='backends 4/5: declare the weight operand pre-rounded (b_is_tf32)')
    a = ap.parse_args()
    if a.what == 'gemm':
        bench_gemm(a.backend, a.only, a.layouts, a.bexact)
    if a.what == 'epi':
        bench_epi(a.backend)
    if a.what == 'misc':
        bench_misc()
    if a.what == 'attn':
        bench_attn()
    if a.what == 'dgrn':
        bench_dgrn()
