"""GPU parity of every C-ABI kernel against the CPU oracle / plain fp32 torch on the same seeded inputs.
Tolerances: fp32 SIMT kernels 1e-4 relative to the tensor's scale (reduction-order noise only)."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import PKG_NAME
from oracle import airnet, freq, uformer

pytestmark = pytest.mark.gpu
import importlib


@pytest.fixture(scope='module')
def ops():
    return importlib.import_module(PKG_NAME + '.ops')


def dev(t):
    return t.cuda().contiguous()


def close(a, b, tol=1e-4, what=''):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(b.abs().max().item(), 1e-6)
    err = (a - b).abs().max().item()
    assert err <= tol * scale + 1e-7, f'{what}: max err {err:.3e} vs scale {scale:.3e}'


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g) * scale


# ----------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (64, 56, 56), (300, 224, 56), (257, 129, 27), (1024, 448, 3584),
                                   (70, 3, 1008), (33, 65, 200)])
@pytest.mark.parametrize('tA,tB', [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_plain(ops, M, N, K, tA, tB):
    A = gen(K, M) if tA else gen(M, K)
    B = gen(N, K, seed=1) if tB else gen(K, N, seed=1)
    C = torch.empty(M, N, device='cuda')
    ops.gemm(dev(A), dev(B), C, transA=tA, transB=tB, backend=1)
    ref = (A.t() if tA else A).double() @ (B.t() if tB else B).double()
    close(C, ref.float(), 2e-5 * math.sqrt(K), f'gemm {M}x{N}x{K}')


def test_gemm_epilogues(ops):
    M, N, K = 192, 224, 56
    A, W, b = gen(M, K), gen(N, K, seed=1), gen(N, seed=2)
    R, aux = gen(M, N, seed=3), gen(M, N, seed=4)
    rs = torch.tensor([0.0, 1.0 / 0.9, 1.0 / 0.9])
    base = A @ W.t() + b
    # bias + GELU + pre-activation output
    C = torch.empty(M, N, device='cuda'); pre = torch.empty(M, N, device='cuda')
    ops.gemm(dev(A), dev(W), C, bias=dev(b), act=ops.ACT_GELU, preact=pre, backend=1)
    close(pre, base, 1e-4, 'preact'); close(C, F.gelu(base), 1e-4, 'gelu')
    # bias + leaky relu
    ops.gemm(dev(A), dev(W), C, bias=dev(b), act=ops.ACT_LRELU, act_param=0.1, backend=1)
    close(C, F.leaky_relu(base, 0.1), 1e-4, 'lrelu')
    # residual + per-sample row scale (DropPath): R + s * (A W^T + b)
    ops.gemm(dev(A), dev(W), C, bias=dev(b), rowscale=dev(rs), rows_per_scale=64, residual=dev(R), backend=1)
    close(C, R + rs.repeat_interleave(64)[:, None] * base, 1e-4, 'residual+rowscale')
    # multiply by GELU'(aux)
    ag = aux.clone().requires_grad_(True)
    F.gelu(ag).sum().backward()
    ops.gemm(dev(A), dev(W), C, aux=dev(aux), aux_act=ops.ACT_GELU, backend=1)
    close(C, (A @ W.t()) * ag.grad, 1e-4, 'dgelu')
    # accumulate + strided output view
    big = torch.zeros(M, 2 * N, device='cuda'); big[:, N:] = dev(R)
    ops.gemm(dev(A), dev(W), big[:, N:], accumulate=True, alpha=0.5, backend=1)
    close(big[:, N:], R + 0.5 * (A @ W.t()), 1e-4, 'accumulate strided')
    assert big[:, :N].abs().max().item() == 0


def test_gemm_splitk_weight_grad(ops):
    M, N, K = 20000, 56, 224          # dW[N,K] = dY^T[N,M] X[M,K], reduction over M
    dY, X = gen(M, N, scale=0.1), gen(M, K, seed=1)
    G0 = gen(N, K, seed=2)
    G = dev(G0)
    ops.gemm(dev(dY), dev(X), G, transA=True, transB=False, accumulate=True, backend=1)
    close(G, G0 + (dY.double().t() @ X.double()).float(), 1e-4, 'split-k')


def test_colsum(ops):
    X = gen(1000, 70)
    rs = torch.rand(10)
    out = torch.empty(70, device='cuda')
    ops.colsum(dev(X), out, rowscale=dev(rs), rows_per_scale=100)
    close(out, (X * rs.repeat_interleave(100)[:, None]).sum(0), 1e-4, 'colsum')
    ops.colsum(dev(X), out, accumulate=True)
    close(out, (X * rs.repeat_interleave(100)[:, None]).sum(0) + X.sum(0), 1e-4, 'colsum acc')


# ----------------------------------------------------------------------------- norms
@pytest.mark.parametrize('rows', [333, 5, 4099])
@pytest.mark.parametrize('C', [28, 56, 112, 224, 448, 768, 896])
def test_layernorm(ops, C, rows):
    x = gen(rows, C).requires_grad_(True)
    g, b = (1 + 0.1 * gen(C, seed=1)).requires_grad_(True), gen(C, seed=2).requires_grad_(True)
    dy, dres = gen(rows, C, seed=3), gen(rows, C, seed=4)
    y = F.layer_norm(x, (C,), g, b)
    y.backward(dy)
    yk, mean, rstd = ops.layernorm_fwd(dev(x.detach()), dev(g.detach()), dev(b.detach()))
    close(yk, y, 1e-5, 'ln fwd')
    dg, db = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    dx = ops.layernorm_bwd(dev(dy), dev(x.detach()), mean, rstd, dev(g.detach()), dev(dres), dg, db)
    close(dx, x.grad + dres, 1e-4, 'ln dx'); close(dg, g.grad, 1e-4, 'ln dgamma'); close(db, b.grad, 1e-4, 'ln dbeta')


def test_batchnorm_head(ops):
    B, C, S = 3, 8, 5000
    x = (gen(B, C, S) * 2 + 0.5).requires_grad_(True)
    w, b = (1 + 0.1 * gen(C, seed=1)).requires_grad_(True), gen(C, seed=2).requires_grad_(True)
    dpool = gen(B, C, seed=3)
    y = F.leaky_relu(F.batch_norm(x, None, None, w, b, training=True, eps=1e-5), 0.1)
    pooled = y.mean(2)
    pooled.backward(dpool)
    xd = dev(x.detach())
    sums = ops.bn_stats(xd, B, C, S)
    n = B * S
    mean = sums[:, 0] / n
    var = sums[:, 1] / n - mean * mean
    close(mean, x.detach().mean((0, 2)), 1e-5, 'bn mean'); close(var, x.detach().var((0, 2), unbiased=False), 1e-4, 'bn var')
    rstd = torch.rsqrt(var + 1e-5).float(); mean = mean.float()
    scale = dev(w.detach()) * rstd
    shift = dev(b.detach()) - mean * scale
    yk, pk = ops.bn_apply(xd, scale, shift, 0.1, B, C, S, want_y=True)
    close(yk, y, 1e-4, 'bn y'); close(pk, pooled, 1e-4, 'bn pooled')
    dx, red = ops.bn_bwd(xd, mean, rstd, scale, shift, 0.1, None, dev(dpool), B, C, S)
    close(dx, x.grad, 1e-3, 'bn dx')
    close(red[:, 1].float(), w.grad, 1e-3, 'bn dweight'); close(red[:, 0].float(), b.grad, 1e-3, 'bn dbias')


# ----------------------------------------------------------------------------- conv pieces
@pytest.mark.parametrize('B,H,W,C', [(2, 16, 16, 40), (1, 10, 70, 136), (3, 8, 8, 4), (1, 64, 64, 224), (5, 33, 17, 12)])
def test_dwconv(ops, B, H, W, C):
    u1 = gen(B, H * W, C).requires_grad_(True)
    w, b = (gen(C, 1, 3, 3, seed=1) * 0.3).requires_grad_(True), gen(C, seed=2).requires_grad_(True)
    dh2 = gen(B, H * W, C, seed=3)
    h1 = F.gelu(u1)
    u2 = F.conv2d(h1.transpose(1, 2).reshape(B, C, H, W), w, b, padding=1, groups=C).flatten(2).transpose(1, 2)
    h2 = F.gelu(u2)
    h2.backward(dh2)
    h1d = dev(h1.detach())
    u2k, h2k = ops.dwconv_fwd(h1d, dev(w.detach()), dev(b.detach()), B, H, W, C)
    close(u2k, u2, 1e-5, 'dw u2'); close(h2k, h2, 1e-5, 'dw h2')
    d2k, h2k2 = ops.dwconv_fwd(h1d, dev(w.detach()), dev(b.detach()), B, H, W, C, u2_mode=1)      # stores gelu'(u2)
    ug = u2.detach().clone().requires_grad_(True)
    F.gelu(ug).sum().backward()
    close(d2k, ug.grad, 1e-5, "dw gelu'(u2)"); assert torch.equal(h2k2, h2k)
    none, h2k3 = ops.dwconv_fwd(h1d, dev(w.detach()), dev(b.detach()), B, H, W, C, u2_mode=None)  # inference: h2 only
    assert none is None and torch.equal(h2k3, h2k)
    du2 = ops.act_bwd(dev(dh2), u2k, ops.ACT_GELU)
    dw, db = torch.zeros(C, 1, 3, 3, device='cuda'), torch.zeros(C, device='cuda')
    du1 = ops.dwconv_bwd(du2, h1d, dev(u1.detach()), dev(w.detach()), dw, db, B, H, W, C)
    close(du1, u1.grad, 1e-4, 'dw du1'); close(dw, w.grad, 1e-4, 'dw dw'); close(db, b.grad, 1e-4, 'dw db')
    dw2, db2 = torch.zeros_like(dw), torch.zeros_like(db)                      # h1 recomputed from u1 inside the kernel
    du1b = ops.dwconv_bwd(du2, None, dev(u1.detach()), dev(w.detach()), dw2, db2, B, H, W, C)
    close(du1b, u1.grad, 1e-4, 'dw du1 (h1 recomputed)'); close(dw2, w.grad, 1e-4, 'dw dw (h1 recomputed)')
    close(db2, b.grad, 1e-4, 'dw db (h1 recomputed)')


@pytest.mark.parametrize('C,k,s,p', [(8, 4, 2, 1), (12, 3, 1, 1), (3, 3, 1, 1)])
def test_im2col_col2im(ops, C, k, s, p):
    B, H, W = 2, 16, 16
    x = gen(B, C, H, W)
    tok = x.flatten(2).transpose(1, 2).contiguous()
    col = ops.im2col(dev(tok), B, H, W, C, k, k, s, p)
    ref = F.unfold(x, k, padding=p, stride=s)                       # [B, C*k*k, L] with (c, ky, kx) order
    L = ref.shape[-1]
    ref = ref.view(B, C, k * k, L).permute(0, 3, 2, 1).reshape(B * L, k * k * C)
    close(col, ref, 1e-6, 'im2col')
    col2 = ops.im2col(dev(x), B, H, W, C, k, k, s, p, nchw_in=True)
    close(col2, ref, 1e-6, 'im2col nchw')
    g = gen(*ref.shape, seed=5)
    dx = ops.col2im(dev(g), B, H, W, C, k, k, s, p)
    gref = g.view(B, L, k * k, C).permute(0, 3, 2, 1).reshape(B, C * k * k, L)
    dref = F.fold(gref, (H, W), k, padding=p, stride=s).flatten(2).transpose(1, 2)
    close(dx, dref, 1e-5, 'col2im')


@pytest.mark.parametrize('C,Co,H,W', [(112, 3, 16, 24), (28, 3, 8, 8), (56, 1, 9, 7)])
def test_conv3x3_out(ops, C, Co, H, W):
    """OutputProj direct conv (tokens -> NCHW image + residual) and its backward vs F.conv2d."""
    B = 2
    t = gen(B, H * W, C).requires_grad_(True)
    w = (gen(Co, C, 3, 3, seed=1) * 0.1).requires_grad_(True)
    b = gen(Co, seed=2).requires_grad_(True)
    ximg = gen(B, Co, H, W, seed=3)
    dout = gen(B, Co, H, W, seed=4)
    ref = F.conv2d(t.view(B, H, W, C).permute(0, 3, 1, 2), w, b, padding=1) + ximg
    ref.backward(dout)
    wk = w.detach().permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    out = ops.conv3x3_out_fwd(dev(t.detach()), dev(wk), dev(b.detach()), dev(ximg), B, H, W, C, Co)
    close(out.view(B, Co, H, W), ref, 1e-5, 'conv3x3_out fwd')
    dt, dW, db = ops.conv3x3_out_bwd(dev(t.detach()), dev(wk), dev(dout), B, H, W, C, Co)
    close(dt, t.grad, 1e-5, 'conv3x3_out dt')
    close(dW.view(Co, 3, 3, C).permute(0, 3, 1, 2), w.grad, 2e-5, 'conv3x3_out dW')
    close(db, b.grad, 2e-5, 'conv3x3_out db')


@pytest.mark.parametrize('B,D,heads,nb', [(3, 448, 16, 3), (2, 448, 1, 3), (5, 100, 4, 2)])
def test_band_coef(ops, B, D, heads, nb):
    """Fused lambda predictor (LN affine -> fc -> Linear, LeakyReLU(0.1), Linear) and its backward vs torch."""
    band = nb - 1
    s = gen(B, D).requires_grad_(True)
    shapes = [(D,), (D,), (heads, D), (heads,), (heads, heads), (heads,), (heads, heads), (heads,)]
    ps = [(gen(*sh, seed=10 + i) * (0.2 if len(sh) == 2 else 1.0)).requires_grad_(True) for i, sh in enumerate(shapes)]
    e = F.linear(s * ps[0] + ps[1], ps[2], ps[3])
    ref = F.linear(F.leaky_relu(F.linear(e, ps[4], ps[5]), 0.1), ps[6], ps[7])
    dout = gen(B, heads, nb, seed=5)
    ref.backward(dout[:, :, band])
    dps = [dev(p.detach()) for p in ps]
    out = torch.zeros(B, heads, nb, device='cuda')
    ops.band_coef_fwd(dev(s.detach()), dps, out, band)
    close(out[:, :, band], ref, 2e-4, 'band_coef fwd')
    assert out[:, :, :band].abs().max().item() == 0
    ds = torch.zeros(B, D, device='cuda')
    init = [gen(*sh, seed=30 + i) for i, sh in enumerate(shapes)]      # kernels accumulate into the gradient buffers
    gs = [dev(t.clone()) for t in init]
    ops.band_coef_bwd(dev(s.detach()), dps, dev(dout), band, ds, gs)
    close(ds, s.grad, 2e-4, 'band_coef dstats')
    for i, (g, p, g0) in enumerate(zip(gs, ps, init)):
        close(g.cpu() - g0, p.grad, 2e-4, f'band_coef grad {i}')


@pytest.mark.parametrize('n,nb,maps', [(128, 4, 6), (16, 2, 5), (64, 8, 3)])
def test_spectral_l1(ops, n, nb, maps):
    """Spectral L1 term (train.py:70,91): l1(D(a), D(b)), D = decompose(inverse=False), vs the oracle under autograd.
    |.| has a sign-discontinuous derivative: a bin whose real or imaginary part is within fp32 FFT round-off of zero
    may flip, which moves the gradient by 2/n of its typical magnitude - so the gradient is held to relative L2 1e-4
    and max error 5 % of the typical magnitude; the loss itself to 1e-5 relative."""
    fd = importlib.import_module(PKG_NAME + '.net.utils.frequency_decompose')
    a = gen(maps, 1, n, n).requires_grad_(True)
    b = gen(maps, 1, n, n, seed=1)
    ref = (freq.decompose(a, 'frequency_decompose', 1. / nb, inverse=False)
           - freq.decompose(b, 'frequency_decompose', 1. / nb, inverse=False)).abs().mean()
    ref.backward()
    bob = fd.half_band_map('frequency_decompose', 1. / nb, n).cuda()
    loss, grad = ops.spectral_l1(dev(a.detach()), dev(b), bob, nb)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    g, gr = grad.cpu().double(), a.grad.double()
    assert (g - gr).norm().item() <= 1e-4 * gr.norm().item() + 2.0 / n * gr.abs().mean().item() * 3
    assert (g - gr).abs().max().item() <= 0.05 * gr.abs().mean().item() * 4
    loss2, none = ops.spectral_l1(dev(a.detach()), dev(b), bob, nb, want_grad=False)
    assert none is None and abs(loss2.item() - loss.item()) <= 1e-6 * abs(loss.item())


def test_pixel_shuffle_and_layout(ops):
    B, H, W, Ci, Co = 2, 4, 4, 16, 8
    x = gen(B, H * W, Ci)
    w, b = gen(Ci, Co, 2, 2, seed=1), gen(Co, seed=2)
    ref = F.conv_transpose2d(x.transpose(1, 2).reshape(B, Ci, H, W), w, b, stride=2).flatten(2).transpose(1, 2)
    wk = w.permute(2, 3, 1, 0).reshape(4 * Co, Ci).contiguous()     # [(ky,kx,co), ci]
    g = torch.empty(B * H * W, 4 * Co, device='cuda')
    ops.gemm(dev(x.view(-1, Ci)), dev(wk), g, bias=dev(b.repeat(4)), backend=1)
    y = torch.zeros(B * 4 * H * W, 2 * Co, device='cuda')
    ops.pixel_shuffle2_fwd(g, y[:, :Co], B, H, W, Co)
    close(y[:, :Co], ref.reshape(-1, Co), 1e-5, 'deconv via gemm+shuffle')
    dg = torch.empty_like(g)
    ops.pixel_shuffle2_bwd(y[:, :Co], dg, B, H, W, Co)
    close(dg, g, 0, 'shuffle adjoint')
    t = gen(B, 100, 3)
    res = gen(B, 3, 100, seed=7)
    close(ops.tokens_to_nchw(dev(t), dev(res), B, 100, 3), t.transpose(1, 2) + res, 1e-6, 'tokens->nchw')
    xx = gen(B, 64, 77)
    close(ops.nchw_to_tokens(dev(xx), B, 77, 64), xx.transpose(1, 2), 0, 'nchw->tokens')
    a, bb = gen(50, 24), gen(50, 24, seed=3)
    dst = torch.zeros(50, 48, device='cuda')
    ops.copy2d(dev(a), dst[:, 24:]); ops.add2d(dev(a), dev(bb), dst[:, :24])
    close(dst, torch.cat([a + bb, a], 1), 0, 'copy/add 2d')


# ----------------------------------------------------------------------------- K1 band filter
def _bob(kind, size, n):
    return freq.band_index_map(kind, size, n, n)[:, :n // 2 + 1].to(torch.uint8).contiguous()


@pytest.mark.parametrize('kind,size,n', [('frequency_decompose_1', 0.5, 64), ('frequency_decompose', 0.25, 64),
                                         ('frequency_decompose_1', 0.5, 128), ('frequency_decompose', 0.125, 32),
                                         ('frequency_decompose_1', 1.0, 16), ('frequency_decompose', 0.5, 8)])
def test_band_split(ops, kind, size, n):
    x = torch.rand(3, 2, n, n, generator=torch.Generator().manual_seed(n))
    nb = freq.num_bands(kind, size)
    y = ops.band_split(dev(x), dev(_bob(kind, size, n)), nb, 0)
    close(y, freq.decompose(x, kind, size, True), 2e-5, f'band_split {kind} {n}')
    ys = ops.band_split(dev(x), dev(_bob(kind, size, n)), nb, 1)          # inverse=False: per-band spectra (re, im)
    close(ys, freq.decompose(x, kind, size, False), 2e-5, f'band spectrum {kind} {n}')


@pytest.mark.parametrize('n', [64, 32])
def test_band_filter_and_energy(ops, n):
    B, nW, heads, nb = 2, 3, 2, 3
    x = torch.rand(B * nW, heads, n, n, generator=torch.Generator().manual_seed(3)).softmax(-1)
    coef = gen(B, heads, nb) * 0.5
    bob = dev(_bob('frequency_decompose_1', 0.5, n))
    y = ops.band_filter(dev(x), bob, dev(coef), nW * heads, heads)
    cf = coef.permute(2, 0, 1).repeat_interleave(nW, 1)                 # [nb, B*nW, heads]
    ref = freq.band_filter(x, 'frequency_decompose_1', 0.5, cf)
    close(y, ref, 2e-5, 'band_filter')
    # self-adjoint and coef gradient
    a = gen(B * nW, heads, n, n, seed=9)
    xg = x.clone().requires_grad_(True); cg = coef.clone().requires_grad_(True)
    (freq.band_filter(xg, 'frequency_decompose_1', 0.5, cg.permute(2, 0, 1).repeat_interleave(nW, 1)) * a).sum().backward()
    close(ops.band_filter(dev(a), bob, dev(coef), nW * heads, heads), xg.grad, 2e-5, 'band_filter adjoint')
    e = torch.zeros(B, heads, nb, device='cuda')
    ops.band_energy(dev(a), dev(x), e, bob, nW * heads, heads)
    close(e, cg.grad, 1e-4, 'band_energy')


def test_dc_split(ops):
    x = gen(5, 3, 64, 64)
    close(ops.dc_split(dev(x)), freq.decompose(x, 'frequency_decompose_dc', 0.5), 1e-5, 'dc_split')


# ----------------------------------------------------------------------------- K2 attention
def ref_win_attn(q, kv, B, H, W, heads, hd, shift, scale, table, coef, nW_img, drop=None):
    C = heads * hd

    def win(t):
        t = t.view(B, H, W, -1)
        if shift:
            t = torch.roll(t, (-shift, -shift), (1, 2))
        return uformer.partition(t).reshape(-1, 64, t.shape[-1])
    qw, kvw = win(q), win(kv)
    qh = qw.view(-1, 64, heads, hd).transpose(1, 2)
    kh = kvw[..., :C].reshape(-1, 64, heads, hd).transpose(1, 2)
    vh = kvw[..., C:].reshape(-1, 64, heads, hd).transpose(1, 2)
    attn = (qh * scale) @ kh.transpose(-1, -2)
    if table is not None:
        attn = attn + table[uformer.rel_index().view(-1)].view(64, 64, heads).permute(2, 0, 1)
    if shift:
        m = uformer.shift_mask(H, W)
        attn = (attn.view(B, -1, heads, 64, 64) + m[None, :, None]).view(-1, heads, 64, 64)
    attn = attn.softmax(-1)
    if coef is not None:
        cf = coef.permute(2, 0, 1).repeat_interleave(nW_img, 1)
        attn = freq.band_filter(attn, 'frequency_decompose_1', 0.5, cf)
    if drop is not None:                     # [windows, heads, 64, 64] of 0 or 1/keep
        attn = attn * drop
    o = (attn @ vh).transpose(1, 2).reshape(-1, 8, 8, C)
    o = uformer.reverse(o, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    return o.reshape(B * H * W, C)


@pytest.mark.parametrize('H,heads,hd,shift,use_coef', [(16, 2, 56, 0, True), (16, 1, 56, 4, True), (24, 2, 28, 4, False),
                                                       (8, 3, 64, 0, True), (16, 2, 56, 4, False)])
def test_win_attn(ops, H, heads, hd, shift, use_coef):
    B, W, C, nb = 2, H, heads * hd, 3
    nW = (H // 8) * (W // 8)
    T = B * H * W
    qkv = gen(T, 3 * C, scale=0.5)
    table = gen(225, heads, seed=1, scale=0.3) if hd != 64 else None
    coef = torch.cat([torch.zeros(B, heads, 1), gen(B, heads, nb - 1, seed=2) * 0.5], -1) if use_coef else None
    dO = gen(T, C, seed=3)
    scale = hd ** -0.5
    leaves = [qkv.clone().requires_grad_(True)]
    tb = table.clone().requires_grad_(True) if table is not None else None
    cf = coef.clone().requires_grad_(True) if coef is not None else None
    ref = ref_win_attn(leaves[0][:, :C], leaves[0][:, C:], B, H, W, heads, hd, shift, scale, tb, cf, nW)
    ref.backward(dO)
    qkvd = dev(qkv)
    bob = dev(_bob('frequency_decompose_1', 0.5, 64))
    o = torch.empty(T, C, device='cuda')
    args = (B, H, W, heads, hd, shift, scale)
    tbd = dev(table) if table is not None else None
    cfd = dev(coef) if coef is not None else None
    ops.win_attn_fwd(qkvd[:, :C], qkvd[:, C:], o, *args, tbd, cfd, heads, bob, nb)
    close(o, ref, 5e-5, 'win_attn fwd')
    dqkv = torch.empty(T, 3 * C, device='cuda')
    dq = torch.empty(T, C, device='cuda'); dkv = torch.empty(T, 2 * C, device='cuda')
    dtab = torch.zeros(225, heads, device='cuda') if table is not None else None
    dcf = torch.zeros(B, heads, nb, device='cuda') if coef is not None else None
    ops.win_attn_bwd(qkvd[:, :C], qkvd[:, C:], dev(dO), dq, dkv, *args, tbd, dtab, cfd, heads, dcf, bob, nb)
    g = leaves[0].grad
    close(dq, g[:, :C], 2e-4, 'win_attn dq'); close(dkv, g[:, C:], 2e-4, 'win_attn dkv')
    if table is not None:
        close(dtab, tb.grad, 2e-4, 'win_attn dtable')
    if coef is not None:
        close(dcf[..., 1:], cf.grad[..., 1:], 5e-4, 'win_attn dcoef')


def drop_mask_host(seed, n_items, p):
    """The kernel's attention-dropout mask (win_attn.cu drop_scale: splitmix64 of seed ^ ((item << 12 | ij) * golden))."""
    import numpy as np
    with np.errstate(over='ignore'):
        idx = (np.arange(n_items, dtype=np.uint64)[:, None] << np.uint64(12)) | np.arange(4096, dtype=np.uint64)[None, :]
        x = np.uint64(seed) ^ (idx * np.uint64(0x9E3779B97F4A7C15))
        x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    u = (x >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    keep = torch.from_numpy((u >= np.float32(p)).astype(np.float32))
    return keep / (1.0 - p)


@pytest.mark.parametrize('H,heads,hd,shift', [(8, 12, 64, 0), (16, 2, 56, 4)])
def test_win_attn_dropout(ops, H, heads, hd, shift):
    """nn.Dropout on the attention map (encoder_ViT.py:94) inside the kernel: the mask is a stateless hash of a device
    seed, so the test rebuilds it on the host and checks forward and backward against the dense oracle with the SAME mask;
    keep-rate within sampling error of 1 - p."""
    B, W, C, nb, p = 2, H, heads * hd, 3, 0.1
    nW = (H // 8) * (W // 8)
    T = B * H * W
    qkv = gen(T, 3 * C, scale=0.5)
    table = gen(225, heads, seed=1, scale=0.3) if hd != 64 else None
    coef = torch.cat([torch.zeros(B, heads, 1), gen(B, heads, nb - 1, seed=2) * 0.5], -1)
    dO = gen(T, C, seed=3)
    scale = hd ** -0.5
    seed = 0x1234567890ABCDE
    mask = drop_mask_host(seed, B * nW * heads, p).view(B * nW, heads, 64, 64)
    assert abs((mask > 0).float().mean().item() - (1 - p)) < 0.01
    leaf = qkv.clone().requires_grad_(True)
    tb = table.clone().requires_grad_(True) if table is not None else None
    cf = coef.clone().requires_grad_(True)
    ref = ref_win_attn(leaf[:, :C], leaf[:, C:], B, H, W, heads, hd, shift, scale, tb, cf, nW, drop=mask)
    ref.backward(dO)
    qkvd = dev(qkv)
    bob = dev(_bob('frequency_decompose_1', 0.5, 64))
    sd = torch.tensor([seed], dtype=torch.int64, device='cuda')
    o = torch.empty(T, C, device='cuda')
    args = (B, H, W, heads, hd, shift, scale)
    tbd = dev(table) if table is not None else None
    ops.win_attn_fwd(qkvd[:, :C], qkvd[:, C:], o, *args, tbd, dev(coef), heads, bob, nb, p, sd)
    close(o, ref, 5e-5, 'win_attn dropout fwd')
    dq = torch.empty(T, C, device='cuda'); dkv = torch.empty(T, 2 * C, device='cuda')
    dtab = torch.zeros(225, heads, device='cuda') if table is not None else None
    dcf = torch.zeros(B, heads, nb, device='cuda')
    ops.win_attn_bwd(qkvd[:, :C], qkvd[:, C:], dev(dO), dq, dkv, *args, tbd, dtab, dev(coef), heads, dcf, bob, nb, p, sd)
    g = leaf.grad
    close(dq, g[:, :C], 2e-4, 'win_attn dropout dq'); close(dkv, g[:, C:], 2e-4, 'win_attn dropout dkv')
    close(dcf[..., 1:], cf.grad[..., 1:], 5e-4, 'win_attn dropout dcoef')
    # another seed gives another mask
    o2 = torch.empty_like(o)
    ops.win_attn_fwd(qkvd[:, :C], qkvd[:, C:], o2, *args, tbd, dev(coef), heads, bob, nb, p, sd + 1)
    assert (o2 - o).abs().max().item() > 1e-3


@pytest.mark.parametrize('H,heads,shift,kind,L', [(16, 2, 0, 'intra', 3), (16, 1, 4, 'inter', 3), (8, 2, 0, 'inter', 2)])
def test_joint_attn(ops, H, heads, shift, kind, L):
    B, W, hd = 2, H, 28
    C = heads * hd
    T = L * B * H * W
    qkv = gen(T, 3 * C, scale=0.5).requires_grad_(True)
    tables = (gen(L * L, 225, heads, seed=1) * 0.3).requires_grad_(True)
    dO = gen(T, C, seed=3)
    # reference: the oracle's joint attention core on the gathered windows
    def win(t):
        t = t.view(L * B, H, W, -1)
        if shift:
            t = torch.roll(t, (-shift, -shift), (1, 2))
        return uformer.partition(t).reshape(-1, 64, t.shape[-1])

    def heads_of(t):
        return t.reshape(t.shape[0], 64, heads, hd).permute(0, 2, 1, 3)
    w = win(qkv)
    mask = uformer.shift_mask(H, W) if shift else None
    a = uformer.joint_attention_core(heads_of(w[..., :C]), heads_of(w[..., C:2 * C]), heads_of(w[..., 2 * C:]),
                                     [tables[i] for i in range(L * L)], mask, L, kind)
    a = a.transpose(1, 2).reshape(-1, 64, C)
    ref = uformer.reverse(a.view(-1, 8, 8, C), H, W)
    if shift:
        ref = torch.roll(ref, (shift, shift), (1, 2))
    ref = ref.reshape(T, C)
    ref.backward(dO)
    qd = dev(qkv.detach())
    o = torch.empty(T, C, device='cuda')
    kd = 0 if kind == 'intra' else 1
    td = dev(tables.detach())
    ops.joint_attn_fwd(qd[:, :C], qd[:, C:], o, L, B, H, W, heads, hd, shift, hd ** -0.5, td, kd)
    close(o, ref, 5e-5, 'joint fwd')
    dq = torch.empty(T, C, device='cuda'); dkv = torch.empty(T, 2 * C, device='cuda')
    dt = torch.zeros_like(td)
    ops.joint_attn_bwd(qd[:, :C], qd[:, C:], dev(dO), dq, dkv, L, B, H, W, heads, hd, shift, hd ** -0.5, td, dt, kd)
    close(dq, qkv.grad[:, :C], 2e-4, 'joint dq'); close(dkv, qkv.grad[:, C:], 2e-4, 'joint dkv')
    close(dt, tables.grad, 2e-4, 'joint dtables')


# ----------------------------------------------------------------------------- DGRN pieces
@pytest.mark.parametrize('B,H,W,C,Co,ldom', [(2, 12, 10, 8, 6, 27), (2, 9, 11, 64, 16, 27), (1, 16, 16, 64, 8, 32)])
def test_dcn(ops, B, H, W, C, Co, ldom):
    """DCNv2 gather / scatter kernels (generic and the pixel-per-warp C = 64 ones; om rows 27 or 32 wide) against the
    oracle's dense restatement through autograd."""
    x = gen(B, C, H, W).requires_grad_(True)
    om = (gen(B, 27, H, W, seed=1) * 1.5).requires_grad_(True)
    w = gen(Co, C, 3, 3, seed=2).requires_grad_(True)
    dout = gen(B, Co, H, W, seed=3)
    o1, o2, m = torch.chunk(om, 3, 1)
    ref = airnet.modulated_deform_conv2d(x, torch.cat((o1, o2), 1), torch.sigmoid(m), w)
    ref.backward(dout)
    xt = dev(x.detach().flatten(2).transpose(1, 2))
    omt = om.detach().flatten(2).transpose(1, 2)
    if ldom > 27:
        omt = torch.cat([omt, torch.full((B, H * W, ldom - 27), 7.0)], -1)          # pad columns must be ignored
    omt = dev(omt.reshape(B * H * W, ldom))
    col = ops.dcn_im2col(xt, omt, B, H, W, C)
    wk = dev(w.detach().permute(0, 2, 3, 1).reshape(Co, 9 * C))
    out = torch.empty(B * H * W, Co, device='cuda')
    ops.gemm(col, wk, out, backend=1)
    close(out.view(B, H * W, Co).transpose(1, 2).reshape(B, Co, H, W), ref, 1e-4, 'dcn fwd')
    dot = dev(dout.flatten(2).transpose(1, 2).reshape(-1, Co))
    dcol = torch.empty_like(col)
    ops.gemm(dot, wk, dcol, transB=False, backend=1)
    dx, dom = ops.dcn_col2im(xt, omt, dcol, B, H, W, C)
    close(dx.transpose(1, 2).reshape(B, C, H, W), x.grad, 2e-4, 'dcn dx')
    close(dom[:, :27].reshape(B, H * W, 27).transpose(1, 2).reshape(B, 27, H, W), om.grad, 2e-3, 'dcn dom')


def test_sft_fuse(ops):
    n = 5000
    x, d, g, b, do = (gen(n, seed=i).requires_grad_(True) for i in range(5))
    out = F.leaky_relu(x + d + x * g + b, 0.1)
    out.backward(do.detach())
    xs = [dev(t.detach()) for t in (x, d, g, b)]
    close(ops.sft_fuse_fwd(*xs, 0.1), out, 1e-6, 'sft fwd')
    grads = ops.sft_fuse_bwd(*xs, dev(do.detach()), 0.1)
    for k, t in zip(grads, (x, d, g, b)):
        close(k, t.grad, 1e-6, 'sft bwd')


# ----------------------------------------------------------------------------- elementwise / optimiser
def test_act_l1_momentum_adam(ops):
    n = 10007
    x, dy = gen(n), gen(n, seed=1)
    for act, fn in ((ops.ACT_GELU, F.gelu), (ops.ACT_LRELU, lambda t: F.leaky_relu(t, 0.1))):
        xr = x.clone().requires_grad_(True)
        fn(xr).backward(dy)
        close(ops.act_fwd(dev(x), act, 0.1), fn(x), 1e-6, 'act fwd'); close(ops.act_bwd(dev(dy), dev(x), act, 0.1), xr.grad, 1e-5, 'act bwd')
    a, b = gen(n, seed=2), gen(n, seed=3)
    loss, grad = ops.l1_loss(dev(a), dev(b))
    close(loss, (a - b).abs().mean().view(1), 1e-5, 'l1'); close(grad, torch.sign(a - b) / n, 1e-6, 'l1 grad')
    k, q = dev(a), dev(b)
    ops.momentum_update(k, q, 0.999)
    close(k, a * 0.999 + b * (1 - 0.999), 1e-6, 'momentum')
    p, g = gen(n, seed=4), gen(n, seed=5) * 0.01
    m, v = torch.zeros(n), torch.zeros(n)
    pd, md, vd = dev(p), dev(m), dev(v)
    for step in (1, 2, 3):
        p, m, v = airnet.adam_step(p, g, m, v, step, 2e-4)
        ops.adam_step(pd, dev(g), md, vd, 2e-4, 0.9, 0.999, 1e-8, step)
    close(pd, p, 1e-6, 'adam p'); close(vd, v, 1e-5, 'adam v')
    # against torch.optim.Adam itself
    pt = torch.nn.Parameter(gen(n, seed=4).clone())
    opt = torch.optim.Adam([pt], lr=2e-4)
    for _ in range(3):
        pt.grad = g.clone()
        opt.step()
    close(pd, pt.detach(), 1e-6, 'adam vs torch.optim')


# ----------------------------------------------------------------------------- training-pair generator (section 8f)
def test_crop_augment_matches_reference_golden(ops):
    """fa_crop_augment against the golden vectors made by the reference's own pipeline functions: bit exact."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'datagen.npz'))
    P = int(g['P'])
    gt, noisy = g['gt'], g['noisy']
    H, W = gt.shape[:2]
    pool = torch.from_numpy(np.concatenate([gt.reshape(-1), noisy.reshape(-1)])).cuda()
    meta = torch.tensor([[0, gt.size, H, W, int(g[f'origin{m}'][0]), int(g[f'origin{m}'][1]), m, 0] for m in range(8)],
                        dtype=torch.int64).cuda()
    deg, clean = ops.crop_augment(pool, meta, torch.zeros(8, device='cuda'), None, P)
    for m in range(8):
        assert np.array_equal(deg[m].cpu().numpy(), g[f'deg{m}']), m
        assert np.array_equal(clean[m].cpu().numpy(), g[f'clean{m}']), m


def test_crop_augment_noise_synthesis(ops):
    """Synthesised degradation: explicit noise -> bit exact against the oracle; stateless generator -> integer side
    exact by construction, Box-Muller to float round-off: at most 1 LSB on < 0.5 % of the pixels; overlapping crops of
    one image share their noise."""
    import numpy as np
    from oracle import datagen as od
    rng = np.random.RandomState(3)
    H, W, P = 48, 64, 16
    img = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    pool = torch.from_numpy(img.reshape(-1)).cuda()
    draws = [(5, 7, 1), (0, 0, 2), (32, 48, 7), (11, 40, 4), (20, 20, 6), (20, 24, 0)]
    meta = torch.tensor([[0, -1, H, W, y, x, m, 99] for y, x, m in draws], dtype=torch.int64).cuda()
    sig = torch.tensor([15., 25., 50., 25., 25., 25.]).cuda()
    noise = torch.from_numpy(rng.randn(len(draws), P, P, 3).astype(np.float32))
    deg, clean = ops.crop_augment(pool, meta, sig, noise.cuda(), P)
    for i, (y, x, m) in enumerate(draws):
        patch = img[y:y + P, x:x + P]
        ref = od.to_tensor(od.augment(od.add_noise(patch, noise[i].numpy(), float(sig[i])), m))
        assert np.array_equal(deg[i].cpu().numpy(), ref), i
        assert np.array_equal(clean[i].cpu().numpy(), od.to_tensor(od.augment(patch, m))), i
    deg2, _ = ops.crop_augment(pool, meta, sig, None, P)
    field = od.normal_field(99, H, W)
    bad = 0
    for i, (y, x, m) in enumerate(draws):
        noisy = od.add_noise(img, field, float(sig[i]))
        ref = od.to_tensor(od.augment(noisy[y:y + P, x:x + P], m))
        d = np.abs(deg2[i].cpu().numpy() - ref) * 255
        assert d.max() <= 1.001, i
        bad += int((d > 0.5).sum())
    assert bad <= 0.005 * len(draws) * P * P * 3
    # samples 4 and 5 (mode 6 / mode 0, origins 4 columns apart) overlap: undo the augmentation and compare
    a = np.rot90(deg2[4].cpu().numpy().transpose(1, 2, 0), k=-3)            # back to patch coordinates
    b = deg2[5].cpu().numpy().transpose(1, 2, 0)
    assert np.array_equal(a[:, 4:], b[:, :P - 4])


def test_device_train_set_mirrors_reference_draws(ops):
    """DeviceTrainSet: same host-side draws as TrainDataset.__getitem__ for a seeded ``random`` (type round-robin,
    per-type shuffle at wrap-around, two crop + augmentation draws per sample), patches cut by the kernel."""
    import random
    import numpy as np
    from oracle import datagen as od
    dg = importlib.import_module(PKG_NAME + '.datagen')
    rng = np.random.RandomState(5)
    imgs = {'denoising_25': [(f'n{i}', rng.randint(0, 256, size=(40 + i, 50, 3)).astype(np.uint8), None) for i in range(3)],
            'deraining': [(f'r{i}', rng.randint(0, 256, size=(37, 45 + i, 3)).astype(np.uint8),
                           rng.randint(0, 256, size=(37, 45 + i, 3)).astype(np.uint8)) for i in range(2)]}
    P = 16
    ds = dg.DeviceTrainSet(['denoising_25', 'deraining'], imgs, patch_size=P, rng=random.Random(17))
    (names, de_ids), d1, d2, c1, c2 = ds.next_batch(6)
    assert de_ids == ['denoising_25', 'deraining'] * 3 and d1.shape == (6, 3, P, P)
    # replay the reference's draw order with the same generator
    r = random.Random(17)
    order = {0: list(range(3)), 1: list(range(2))}
    it = [0, 0]
    for i in range(6):
        t = i % 2
        if it[t] == 0:
            for k in reversed(range(1, len(order[t]))):
                j = r.randrange(1, k + 1)
                order[t][k], order[t][j] = order[t][j], order[t][k]
        idx = order[t][it[t]]
        name, clean, degraded = imgs[de_ids[i]][idx]
        assert names[i] == name
        clean = dg.crop_img(clean, 16)
        for v, (dv, cv) in enumerate(((d1, c1), (d2, c2))):
            y0 = r.randint(0, clean.shape[0] - P); x0 = r.randint(0, clean.shape[1] - P); mode = r.randint(1, 7)
            assert ds.last_meta[v * 6 + i, 4:7].tolist() == [y0, x0, mode]
            assert np.array_equal(cv[i].cpu().numpy(), od.to_tensor(od.augment(clean[y0:y0 + P, x0:x0 + P], mode)))
            if degraded is not None:
                dimg = dg.crop_img(degraded, 16)
                assert np.array_equal(dv[i].cpu().numpy(), od.to_tensor(od.augment(dimg[y0:y0 + P, x0:x0 + P], mode)))
            else:
                assert not torch.equal(dv[i], cv[i]) and (dv[i] - cv[i]).abs().mean().item() < 0.15
        it[t] = (it[t] + 1) % len(order[t])


@pytest.mark.parametrize('B,H,W,Ci,Co', [(2, 128, 128, 64, 64), (3, 32, 32, 32, 48), (1, 8, 256, 64, 24), (2, 16, 64, 96, 64)])
def test_conv3x3_implicit_gemm(ops, B, H, W, Ci, Co):
    """3x3 s1 p1 convolution on tokens as an implicit GEMM (4-D TMA patch operand, zero padding by out-of-bounds fill):
    forward with bias + LeakyReLU, data gradient through the flipped / transposed weight, weight and bias gradients -
    against fp64 torch conv2d autograd.  Geometries: one image row per M tile (W = 128), several rows per tile (W < 128),
    a row segment per tile (W > 128)."""
    assert ops.conv3x3_eligible(H, W, Ci, Co)
    x = gen(B, H * W, Ci, scale=0.7)
    w = gen(Co, Ci, 3, 3, seed=1, scale=0.1)
    b = gen(Co, seed=2, scale=0.3)
    g = gen(B * H * W, Co, seed=3)
    xr = x.double().view(B, H, W, Ci).permute(0, 3, 1, 2).requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    pre = F.conv2d(xr, wr, br, padding=1)
    yr = F.leaky_relu(pre, 0.1)
    yr.backward(g.double().view(B, H, W, Co).permute(0, 3, 1, 2))
    wk = w.permute(0, 2, 3, 1).reshape(Co, 9 * Ci).contiguous()
    xd, wkd = dev(x), dev(wk)
    y = torch.empty(B * H * W, Co, device='cuda')
    ops.conv3x3_gemm(xd, wkd, y, B, H, W, bias=dev(b), act=ops.ACT_LRELU, act_param=0.1)
    ref = yr.permute(0, 2, 3, 1).reshape(B * H * W, Co)
    close(y, ref, 2e-5, 'conv3x3 implicit fwd')
    # gradient w.r.t. the pre-activation, then the two contractions of the backward
    # (LeakyReLU mask from the fp64 pre-activation: a value within round-off of 0 may change sign between implementations)
    mask = torch.where(pre.detach() > 0, 1.0, 0.1).permute(0, 2, 3, 1).reshape(B * H * W, Co).float()
    gp = dev(g * mask)
    dwk = torch.zeros(Co, 9 * Ci, device='cuda')
    db = torch.zeros(Co, device='cuda')
    ops.conv3x3_wgrad(gp, xd, dwk, B, H, W, accumulate=True, dbias=db)
    close(dwk, wr.grad.permute(0, 2, 3, 1).reshape(Co, 9 * Ci), 5e-5, 'conv3x3 implicit dW')
    close(db, br.grad, 5e-5, 'conv3x3 implicit db')
    if Co % 32 == 0:
        wflip = wk.view(Co, 9, Ci).flip(1).permute(2, 1, 0).reshape(Ci, 9 * Co).contiguous()
        dx = torch.empty(B * H * W, Ci, device='cuda')
        ops.conv3x3_gemm(gp.view(B, H * W, Co), dev(wflip), dx, B, H, W)
        close(dx, xr.grad.permute(0, 2, 3, 1).reshape(B * H * W, Ci), 5e-5, 'conv3x3 implicit dX')
