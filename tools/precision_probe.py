#!/usr/bin/env python
"""CPU experiment behind DESIGN.md's contraction-precision table: the oracle's golden train step (batch 2,
Uformer+Uformer all_3_bands) with every dense contraction (nn.Linear, dense conv, transposed conv and their
backward contractions) evaluated as the tensor core would evaluate it under a given operand treatment, compared
with the same step in plain fp32.

  mode 'rn1'    both operands rounded to TF32 (round-to-nearest, ties away = cvt.rna)      -> 1 MMA per product
  mode 'aexact' first operand exact (hi + lo split), second operand RN-rounded to TF32        -> 2 MMAs per product
  mode 'trunc1' both operands truncated to TF32 (what kind::tf32 does to raw fp32 operands)   -> 1 MMA per product

"first operand" follows fa_gemm's A: the activation in y = x W^T, the output gradient in dX = dY W and in
dW = dY^T X.  Accumulation stays fp32 (the kernel promotes TMEM partial sums to fp32 registers every KC elements).
A class filter restricts the treatment to some layer classes (--only leff,attn,head,conv); the rest stay fp32.

Usage: python tools/precision_probe.py [--modes rn1,aexact,trunc1] [--only leff,attn]
"""
import argparse
import importlib
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
PKG = 'frequency-wised_all-in-one_image_restoration_model_b200'

_linear, _conv2d, _convT = F.linear, F.conv2d, F.conv_transpose2d


def rn(x):
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def tr(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


MODES = {'rn1': (rn, rn), 'aexact': (lambda x: x, rn), 'trunc1': (tr, tr), 'fp32': (lambda x: x, lambda x: x)}
STATE = {'opA': None, 'opB': None, 'only': None, 'cls': None, 'bwd_only': False, 'dw_only': False}


def _ops():
    if STATE['only'] is not None and STATE['cls'] not in STATE['only']:
        return MODES['fp32']
    return STATE['opA'], STATE['opB']


class LinFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        opA, opB = _ops()
        ctx.ops = (opA, opB)
        ctx.save_for_backward(x, W)
        ctx.has_b = b is not None
        if STATE['bwd_only'] or STATE['dw_only']:
            return _linear(x, W, b)
        return _linear(opA(x), opB(W), b)

    @staticmethod
    def backward(ctx, g):
        opA, opB = ctx.ops
        x, W = ctx.saved_tensors
        g2, x2 = g.reshape(-1, g.shape[-1]), x.reshape(-1, x.shape[-1])
        dx = ((g2 @ W) if STATE['dw_only'] else (opA(g2) @ opB(W))).view(x.shape)
        dW = opA(g2).t() @ opB(x2)
        return dx, dW, (g2.sum(0) if ctx.has_b else None)


class ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, stride, padding):
        opA, opB = _ops()
        ctx.ops, ctx.cfg = (opA, opB), (stride, padding)
        ctx.save_for_backward(x, W)
        ctx.has_b = b is not None
        if STATE['bwd_only'] or STATE['dw_only']:
            return _conv2d(x, W, b, stride=stride, padding=padding)
        return _conv2d(opA(x), opB(W), b, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, g):
        opA, opB = ctx.ops
        x, W = ctx.saved_tensors
        s, p = ctx.cfg
        if STATE['dw_only']:
            dx = torch.nn.grad.conv2d_input(x.shape, W, g, stride=s, padding=p)
        else:
            dx = torch.nn.grad.conv2d_input(x.shape, opB(W), opA(g), stride=s, padding=p)
        dW = torch.nn.grad.conv2d_weight(opB(x), W.shape, opA(g), stride=s, padding=p)
        return dx, dW, (g.sum((0, 2, 3)) if ctx.has_b else None), None, None


class ConvTFn(torch.autograd.Function):
    """ConvTranspose2d k2 s2 = a GEMM x[T,Cin] . W[Cin, Co*4] + pixel shuffle."""
    @staticmethod
    def forward(ctx, x, W, b):
        opA, opB = _ops()
        ctx.ops = (opA, opB)
        ctx.save_for_backward(x, W)
        if STATE['bwd_only'] or STATE['dw_only']:
            return _convT(x, W, b, stride=2)
        return _convT(opA(x), opB(W), b, stride=2)

    @staticmethod
    def backward(ctx, g):
        opA, opB = ctx.ops
        x, W = ctx.saved_tensors
        dx = _conv2d(g, W, stride=2) if STATE['dw_only'] else _conv2d(opA(g), opB(W), stride=2)
        # dW[ci,co,ky,kx] = sum x[b,ci,y,x] g[b,co,2y+ky,2x+kx]
        B, Ci, H, Wd = x.shape
        gg = opA(g).view(B, -1, H, 2, Wd, 2)
        dW = torch.einsum('bchw,bohywx->coyx', opB(x), gg)
        return dx, dW, g.sum((0, 2, 3))


def linear(x, W, b=None):
    return LinFn.apply(x, W, b)


def conv2d(x, W, b=None, stride=1, padding=0, dilation=1, groups=1):
    if groups != 1:
        return _conv2d(x, W, b, stride, padding, dilation, groups)
    return ConvFn.apply(x, W, b, stride, padding)


def conv_transpose2d(x, W, b=None, stride=1, **kw):
    assert stride == 2 and not kw
    return ConvTFn.apply(x, W, b)


def run(mode, only):
    from oracle import airnet, detfill, uformer
    from conftest import load_golden, load_spec, t
    synth = importlib.import_module(PKG + '.synth')
    g = load_golden('airnet_uu_train.npz')
    sd = detfill.make_state(load_spec('spec_airnet_uformer_uformer_L3.json'))
    pnames = [k[len('E.E.encoder_q.'):] for k in sd if k.startswith('E.E.encoder_q.')
              and not any(s in k for s in ('running_', 'num_batches', 'relative_position_index', 'mask_freq'))]
    grads_on = [k for k in sd if sd[k].is_floating_point() and not k.startswith('E.E.encoder_k.')
                and not any(s in k for s in ('running_', 'queue', 'mask_freq'))]
    for k in grads_on:
        sd[k].requires_grad_(True)
    dp = {}
    for k in g:
        if k.startswith('dp/'):
            _, blk, i = k.split('/')
            dp.setdefault(blk, [None, None])[int(i)] = t(g[k])
    xq, xk, clean = synth.noisy_batch(2, 25)
    STATE['opA'], STATE['opB'] = MODES[mode]
    STATE['only'] = only

    # layer classes by call site: patch the oracle's helpers so each contraction knows its class
    orig = dict(lin=uformer.lin, leff=uformer.leff)

    def lin_cls(sd_, p, x):
        prev = STATE['cls']
        if STATE['cls'] is None:
            STATE['cls'] = 'head' if ('mlp_head' in p or '.mlp.' in p or p.endswith('mlp.0') or p.endswith('mlp.2')) else 'attn'
        try:
            return orig['lin'](sd_, p, x)
        finally:
            STATE['cls'] = prev

    def leff_cls(sd_, p, x):
        STATE['cls'] = 'leff'
        try:
            return orig['leff'](sd_, p, x)
        finally:
            STATE['cls'] = None

    uformer.lin, uformer.leff = lin_cls, leff_cls
    airnet.lin = lin_cls
    F.linear, F.conv2d, F.conv_transpose2d = linear, lambda *a, **k: _with_cls('conv', conv2d, *a, **k), \
        lambda *a, **k: _with_cls('conv', conv_transpose2d, *a, **k)
    try:
        restored, logits, kout = airnet.airnet_uformer_forward(sd, xq, xk, True, dp=dp, param_names=pnames)
        labels = torch.zeros(2, dtype=torch.long)
        ce = sum(torch.nn.functional.cross_entropy(l, labels) for l in logits) / 3
        loss = (restored - clean).abs().mean() + 0.6 * ce
        loss.backward()
    finally:
        F.linear, F.conv2d, F.conv_transpose2d = _linear, _conv2d, _convT
        uformer.lin, uformer.leff = orig['lin'], orig['leff']
        airnet.lin = orig['lin']
    grads = {k: sd[k].grad.detach().clone() for k in grads_on if sd[k].grad is not None}
    return dict(restored=restored.detach(), logits=torch.stack(logits).detach(), loss=loss.item(), grads=grads)


def _with_cls(cls, fn, *a, **k):
    prev = STATE['cls']
    if prev is None:
        STATE['cls'] = cls
    try:
        return fn(*a, **k)
    finally:
        STATE['cls'] = prev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--modes', default='rn1,aexact,trunc1')
    ap.add_argument('--only', default=None, help='comma list of layer classes to treat (leff, attn, head, conv); default all')
    ap.add_argument('--json', default=None)
    ap.add_argument('--bwd-only', action='store_true', help='forward contractions stay fp32; only dX and dW get the treatment')
    ap.add_argument('--dw-only', action='store_true', help='only the weight-gradient contractions get the treatment')
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    base = run('fp32', None)
    STATE['bwd_only'], STATE['dw_only'] = a.bwd_only, a.dw_only
    only = set(a.only.split(',')) if a.only else None
    report = {}
    for mode in a.modes.split(','):
        r = run(mode, only)
        worst_rel, worst_abs, worst_name, n_over = 0.0, 0.0, None, 0
        for k, gref in base['grads'].items():
            scale = max(gref.abs().max().item(), 1e-12)
            err = (r['grads'][k] - gref).abs().max().item()
            if err / scale > worst_rel:
                worst_rel, worst_name = err / scale, k
            worst_abs = max(worst_abs, err)
            n_over += int(err > 1e-3 * max(scale, 1.0) and err > 1e-3)
        rep = {'restored_maxabs': (r['restored'] - base['restored']).abs().max().item(),
               'logits_maxabs': (r['logits'] - base['logits']).abs().max().item(),
               'loss_abs': abs(r['loss'] - base['loss']),
               'grad_maxabs': worst_abs, 'grad_worst_rel_to_own_scale': worst_rel, 'grad_worst_name': worst_name,
               'grads_over_1e-3_abs': n_over}
        report[mode] = rep
        print(mode, 'only=' + str(a.only), 'bwd_only' if a.bwd_only else ('dw_only' if a.dw_only else ''), json.dumps(rep))
    if a.json:
        with open(a.json, 'w') as f:
            json.dump({'only': a.only, 'report': report}, f, indent=1)


if __name__ == '__main__':
    main()
