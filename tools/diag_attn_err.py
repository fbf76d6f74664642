"""Accuracy probe: fused window attention (fwd/bwd) against an fp64 evaluation of the same math."""
import importlib, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_ops as T
ops = importlib.import_module(T.PKG_NAME + '.ops')
for (H, heads, hd, shift, use_coef, qs) in [(8, 3, 64, 0, True, 0.5), (8, 3, 64, 0, True, 2.0), (16, 2, 56, 4, True, 0.5), (16, 2, 28, 0, False, 0.5)]:
    B, W, C, nb = 2, H, heads * hd, 3
    nW = (H // 8) * (W // 8); Tn = B * H * W
    qkv = T.gen(Tn, 3 * C, scale=qs)
    table = T.gen(225, heads, seed=1, scale=0.3) if hd != 64 else None
    coef = torch.cat([torch.zeros(B, heads, 1), T.gen(B, heads, nb - 1, seed=2) * 0.5], -1) if use_coef else None
    dO = T.gen(Tn, C, seed=3)
    scale = hd ** -0.5
    outs = {}
    for dt in (torch.float64, torch.float32):
        leaf = qkv.clone().to(dt).requires_grad_(True)
        tb = table.clone().to(dt) if table is not None else None
        cf = coef.clone().to(dt) if coef is not None else None
        ref = T.ref_win_attn(leaf[:, :C], leaf[:, C:], B, H, W, heads, hd, shift, scale, tb, cf, nW)
        ref.backward(dO.to(dt))
        outs[dt] = (ref.detach().double(), leaf.grad.double())
    qkvd = qkv.cuda()
    bob = T._bob('frequency_decompose_1', 0.5, 64).cuda()
    o = torch.empty(Tn, C, device='cuda')
    args = (B, H, W, heads, hd, shift, scale)
    tbd = table.cuda() if table is not None else None
    cfd = coef.cuda() if coef is not None else None
    ops.win_attn_fwd(qkvd[:, :C], qkvd[:, C:], o, *args, tbd, cfd, heads, bob, nb)
    dq = torch.empty(Tn, C, device='cuda'); dkv = torch.empty(Tn, 2 * C, device='cuda')
    dtab = torch.zeros(225, heads, device='cuda') if table is not None else None
    dcf = torch.zeros(B, heads, nb, device='cuda') if coef is not None else None
    ops.win_attn_bwd(qkvd[:, :C], qkvd[:, C:], dO.cuda(), dq, dkv, *args, tbd, dtab, cfd, heads, dcf, bob, nb)
    g = torch.cat([dq, dkv], 1).double().cpu()
    r64o, r64g = outs[torch.float64]; r32o, r32g = outs[torch.float32]
    def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
    print(f'H{H} heads{heads} hd{hd} shift{shift} coef{use_coef} qscale{qs}: fwd ours {rel(o.double().cpu(), r64o):.2e} cpu32 {rel(r32o, r64o):.2e} | '
          f'bwd ours {rel(g, r64g):.2e} cpu32 {rel(r32g, r64g):.2e}')
