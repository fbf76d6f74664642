"""Thin torch-tensor front end of the C ABI (``include/freqair.h``).

Every function here only validates arguments, pulls raw pointers out of torch tensors and enqueues
kernels on ``torch.cuda.current_stream()``.  PyTorch is the allocator and the stream owner; all
arithmetic happens in ``libfreqair.so``.  Nothing here can run without a CUDA device.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import FaGemmEpilogue

ACT_NONE, ACT_GELU, ACT_LRELU, ACT_SIGMOID = 0, 1, 2, 3
ACT_MUL = 4                     # aux_act only: aux already holds the derivative, multiply by it
FLOP_COUNTER = [None]          # bench.py roofline leg: set to 0 to accumulate 2*M*N*K of every fa_gemm call
BYTE_COUNTER = [None]          # bench.py HBM roofline leg: set to 0 to accumulate the ALGORITHMIC bytes
                               # of every depthwise-conv / LayerNorm launch
# fa_gemm backend used when a call does not name one: 0 = auto (tcgen05 3xTF32 where eligible, else fp32 SIMT).
# FREQAIR_GEMM_BACKEND=1 forces the fp32 SIMT kernel everywhere (A/B accuracy and speed comparisons).
DEFAULT_GEMM_BACKEND = int(os.environ.get('FREQAIR_GEMM_BACKEND', '0'))
GEMM_3X, GEMM_SIMT, GEMM_2X, GEMM_1X = 0, 1, 4, 5     # fa_gemm backend ids (include/freqair.h)
# Contractions of the restorer's LeFF class (linear1, linear2 and their backward contractions: 62 % of the decoder's
# flops) run 1xTF32 with both operands rounded to nearest: measured on the golden train step (tools/precision_probe.py
# on the CPU, tests/test_gpu_golden.py on the GPU; table in DESIGN.md section 3) this moves the restored image by <= 5e-4
# and no gradient by more than 1.5e-4 - inside the 1e-3 bar.  Every other layer class (attention projections,
# convolutions, encoder heads) exceeds or touches the bar under TF32 rounding and stays on the error-compensated 3xTF32
# path, and so does the ENCODER's LeFF: its bottleneck tokens (`inter`) moved by 1.14e-3 under 1xTF32.
# FREQAIR_LEFF_BACKEND / FREQAIR_LEFF_ENC_BACKEND select the fa_gemm backend of the two (0 = 3xTF32, 4 = 2x, 5 = 1x).
# BACKWARD contractions (dX = dY.W and dW = dY^T.X) of the Uformer path - every layer class - also run 1xTF32: they do
# not touch the forward activations at all, and on the golden train step rounding both operands of every backward
# contraction moves no gradient by more than 1.1e-4 (tools/precision_probe.py --bwd-only; bar 1e-3).  The ViT / ResNet /
# DGRN paths keep 3xTF32 in both directions (ill-conditioned 12-layer ViT, discontinuous DCN offset gradients).
# FREQAIR_BWD_BACKEND=0 puts them back on 3xTF32.
LEFF_BACKEND = int(os.environ.get('FREQAIR_LEFF_BACKEND', str(GEMM_1X)))
BWD_BACKEND = int(os.environ.get('FREQAIR_BWD_BACKEND', str(GEMM_1X)))
LEFF_ENC_BACKEND = int(os.environ.get('FREQAIR_LEFF_ENC_BACKEND', str(GEMM_3X)))
K_GEMM, K_WIN_ATTN, K_JOINT_ATTN, K_BAND, K_LN, K_DWCONV, K_IM2COL, K_BN, K_OPTIM, K_DCN, K_ELEM = range(1, 12)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('freqair ops need CUDA tensors: there is no CPU path')
    return ctypes.c_void_p(t.data_ptr())


def _f32(*ts):
    for t in ts:
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise RuntimeError(f'freqair: expected a contiguous float32 tensor, got {t.dtype} strides {t.stride()}')


PROFILE = None                 # tools/profile_step.py: list that receives (name, int-args signature, event0, event1)


def _call(name, *args):
    lib = _lib.load()
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(getattr(lib, name)(*args), name)
        e1.record()
        PROFILE.append((name, tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool) and abs(a) < (1 << 31)), e0, e1))
        return
    _lib.check(getattr(lib, name)(*args), name)


def launch_count(reset=False):
    lib = _lib.load()
    n = lib.fa_launch_count()
    if reset:
        lib.fa_launch_count_reset()
    return int(n)


def prof_begin(kernel_class):
    _call('fa_prof_begin', int(kernel_class))


def prof_end():
    ms, n = ctypes.c_double(0), ctypes.c_int64(0)
    _call('fa_prof_end', ctypes.byref(ms), ctypes.byref(n))
    return ms.value, n.value


# ----------------------------------------------------------------------------- GEMM
def _rows2d(t):
    """(ptr, rows, cols, ld) of a 2-D row-major view whose last dim is dense."""
    if t.dim() != 2 or t.stride(1) != 1 or t.dtype != torch.float32:
        raise RuntimeError(f'freqair.gemm: need a 2-D float32 row-major view, got shape {tuple(t.shape)} strides {t.stride()}')
    return t.shape[0], t.shape[1], (t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1]))


def gemm(A, B, C, transA=False, transB=True, bias=None, act=ACT_NONE, act_param=0.0, aux=None, aux_act=ACT_NONE,
         aux_param=0.0, rowscale=None, rows_per_scale=1, residual=None, accumulate=False, alpha=1.0, preact=None,
         backend=None, a_rowsum=None, a_kscale=None, a_k_rows_per_scale=1, b_is_tf32=False):
    """C = epi(alpha * op(A) @ op(B)); see fa_gemm in include/freqair.h.  A, B, C, aux, residual, preact are 2-D
    row-major views (row stride may exceed the width).  transB=True means B is an nn.Linear weight [N, K]."""
    ar, ac, lda = _rows2d(A)
    br, bc, ldb = _rows2d(B)
    M, K = (ac, ar) if transA else (ar, ac)
    Kb, N = (bc, br) if transB else (br, bc)
    if K != Kb:
        raise RuntimeError(f'freqair.gemm: inner dims differ ({K} vs {Kb})')
    cr, cc, ldc = _rows2d(C)
    if (cr, cc) != (M, N):
        raise RuntimeError(f'freqair.gemm: C is {cr}x{cc}, expected {M}x{N}')
    if FLOP_COUNTER[0] is not None:
        FLOP_COUNTER[0] += 2 * M * N * K
    if backend is None:
        backend = DEFAULT_GEMM_BACKEND
    e = FaGemmEpilogue()
    e.bias = bias.data_ptr() if bias is not None else None
    e.act, e.act_param = act, act_param
    if aux is not None:
        _, _, e.ldaux = _rows2d(aux)
        e.aux = aux.data_ptr()
    e.aux_act, e.aux_param = aux_act, aux_param
    if rowscale is not None:
        e.rowscale = rowscale.data_ptr()
    e.rows_per_scale = rows_per_scale
    if residual is not None:
        _, _, e.ldr = _rows2d(residual)
        e.residual = residual.data_ptr()
    e.accumulate = 1 if accumulate else 0
    e.alpha = alpha
    if preact is not None:
        _, _, e.ldpre = _rows2d(preact)
        e.preact = preact.data_ptr()
    if a_rowsum is not None:            # accumulated: a_rowsum[m] += sum_k op(A)[m, k]  (bias gradient of a transA GEMM)
        _f32(a_rowsum)
        if a_rowsum.numel() != M:
            raise RuntimeError(f'freqair.gemm: a_rowsum has {a_rowsum.numel()} entries, expected M={M}')
        e.a_rowsum = a_rowsum.data_ptr()
    if a_kscale is not None:            # op(A)[m, k] *= a_kscale[k // a_k_rows_per_scale] inside the contraction
        _f32(a_kscale)
        if a_kscale.numel() * a_k_rows_per_scale < K:
            raise RuntimeError(f'freqair.gemm: a_kscale covers {a_kscale.numel() * a_k_rows_per_scale} of K={K}')
        e.a_kscale = a_kscale.data_ptr()
        e.a_k_rows_per_scale = a_k_rows_per_scale
    e.b_is_tf32 = 1 if b_is_tf32 else 0
    _call('fa_gemm', _p(A), _p(B), _p(C), M, N, K, lda, ldb, ldc, int(transA), int(transB), ctypes.byref(e), backend,
          _stream())
    return C


def conv3x3_eligible(H, W, Cin, Cout=8):
    """Geometry the implicit-GEMM 3x3 convolution accepts (fa_conv3x3_gemm / fa_conv3x3_wgrad)."""
    return (Cin % 32 == 0 and Cout >= 8 and Cout % 4 == 0 and (H * W) % 128 == 0 and W % 32 == 0
            and (W % 128 == 0 or 128 % W == 0))


def conv3x3_gemm(x, wk, y, B, H, W, bias=None, act=ACT_NONE, act_param=0.0, residual=None, aux=None, aux_act=ACT_NONE,
                 aux_param=0.0, accumulate=False, backend=0):
    """y[T, Cout] = epi(conv3x3_s1_p1(x [B,H*W,Cin]) with wk [Cout, 9*Cin]) without a patch matrix (implicit GEMM)."""
    _f32(x, wk)
    Cin, Cout = x.shape[-1], wk.shape[0]
    _, _, ldy = _rows2d(y)
    if FLOP_COUNTER[0] is not None:
        FLOP_COUNTER[0] += 2 * B * H * W * Cout * 9 * Cin
    e = FaGemmEpilogue()
    e.bias = bias.data_ptr() if bias is not None else None
    e.act, e.act_param, e.alpha = act, act_param, 1.0
    e.rows_per_scale = 1
    if residual is not None:
        _, _, e.ldr = _rows2d(residual)
        e.residual = residual.data_ptr()
    if aux is not None:
        _, _, e.ldaux = _rows2d(aux)
        e.aux = aux.data_ptr()
    e.aux_act, e.aux_param = aux_act, aux_param
    e.accumulate = 1 if accumulate else 0
    _call('fa_conv3x3_gemm', _p(x), _p(wk), _p(y), B, H, W, Cin, Cout, ldy, ctypes.byref(e), backend, _stream())
    return y


def conv3x3_wgrad(g, x, dwk, B, H, W, accumulate=True, dbias=None, backend=0):
    """dwk[Cout, 9*Cin] (+)= g[T, Cout]^T . patches(x); dbias += column sums of g (implicit GEMM, no patch matrix)."""
    _f32(x, dwk)
    Cin = x.shape[-1]
    _, Cout, ldg = _rows2d(g)
    if FLOP_COUNTER[0] is not None:
        FLOP_COUNTER[0] += 2 * B * H * W * Cout * 9 * Cin
    _call('fa_conv3x3_wgrad', _p(g), ldg, _p(x), _p(dwk), B, H, W, Cin, Cout, int(accumulate), _p(dbias), backend, _stream())
    return dwk


def colsum(X, out, rowscale=None, rows_per_scale=1, accumulate=False):
    M, N, ld = _rows2d(X)
    _call('fa_colsum', _p(X), _p(out), M, N, ld, _p(rowscale), rows_per_scale, int(accumulate), _stream())
    return out


# ----------------------------------------------------------------------------- band filter (K1)
def band_split(x, band_of_bin, nbands, mode=0):
    _f32(x)
    n = x.shape[-1]
    nmaps = x.numel() // (n * n)
    shape = (nbands,) + tuple(x.shape) + ((2,) if mode == 1 else ())
    y = torch.empty(shape, device=x.device, dtype=torch.float32)
    _call('fa_band_split', _p(x), _p(y), nmaps, n, _p(band_of_bin), nbands, mode, _stream())
    return y


def band_filter(x, band_of_bin, coef, maps_per_group, heads):
    _f32(x, coef)
    n = x.shape[-1]
    nmaps = x.numel() // (n * n)
    y = torch.empty_like(x)
    _call('fa_band_filter', _p(x), _p(y), nmaps, n, _p(band_of_bin), coef.shape[-1], _p(coef), maps_per_group, heads,
          _stream())
    return y


def band_energy(a, b, out, band_of_bin, maps_per_group, heads):
    _f32(a, b, out)
    n = a.shape[-1]
    nmaps = a.numel() // (n * n)
    _call('fa_band_energy', _p(a), _p(b), _p(out), nmaps, n, _p(band_of_bin), out.shape[-1], maps_per_group, heads,
          _stream())
    return out


def dc_split(x):
    _f32(x)
    n = x.shape[-1]
    y = torch.empty((2,) + tuple(x.shape), device=x.device, dtype=torch.float32)
    _call('fa_dc_split', _p(x), _p(y), x.numel() // (n * n), n, _stream())
    return y


# ----------------------------------------------------------------------------- attention (K2)
def win_attn_fwd(q, kv, o, B, H, W, heads, hd, shift, scale, table, coef, coef_bstride, band_of_bin, nbands, drop_p=0.0,
                 drop_seed=None):
    """drop_seed: int64 CUDA tensor [1] (attention-map dropout with probability drop_p, mask = hash of the seed)."""
    _call('fa_win_attn_fwd', _p(q), q.stride(0), _p(kv), kv.stride(0), _p(o), B, H, W, heads, hd, shift, scale,
          _p(table), _p(coef), coef_bstride, _p(band_of_bin), nbands, drop_p, _p(drop_seed), _stream())


def win_attn_bwd(q, kv, dout, dq, dkv, B, H, W, heads, hd, shift, scale, table, dtable, coef, coef_bstride, dcoef,
                 band_of_bin, nbands, drop_p=0.0, drop_seed=None):
    _f32(dout, dq, dkv)
    _call('fa_win_attn_bwd', _p(q), q.stride(0), _p(kv), kv.stride(0), _p(dout), _p(dq), _p(dkv), B, H, W, heads, hd,
          shift, scale, _p(table), _p(dtable), _p(coef), coef_bstride, _p(dcoef), _p(band_of_bin), nbands, drop_p,
          _p(drop_seed), _stream())


def joint_attn_fwd(q, kv, o, L, B, H, W, heads, hd, shift, scale, tables, kind):
    _call('fa_joint_attn_fwd', _p(q), q.stride(0), _p(kv), kv.stride(0), _p(o), L, B, H, W, heads, hd, shift, scale,
          _p(tables), kind, _stream())


def joint_attn_bwd(q, kv, dout, dq, dkv, L, B, H, W, heads, hd, shift, scale, tables, dtables, kind):
    _f32(dout, dq, dkv)
    _call('fa_joint_attn_bwd', _p(q), q.stride(0), _p(kv), kv.stride(0), _p(dout), _p(dq), _p(dkv), L, B, H, W, heads,
          hd, shift, scale, _p(tables), _p(dtables), kind, _stream())


# ----------------------------------------------------------------------------- norms
def layernorm_fwd(x, gamma, beta, y=None, want_stats=True):
    _f32(x, gamma, beta)
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty_like(x) if y is None else y
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    if BYTE_COUNTER[0] is not None:
        BYTE_COUNTER[0] += 8 * rows * C
    _call('fa_layernorm_fwd', _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), rows, C, _stream())
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dres, dgamma, dbeta, dx=None):
    _f32(dy, x, dres)
    C = x.shape[-1]
    rows = x.numel() // C
    dx = torch.empty_like(x) if dx is None else dx
    if BYTE_COUNTER[0] is not None:            # read dy, x (+ dres), write dx
        BYTE_COUNTER[0] += 4 * rows * C * (3 + (dres is not None))
    _call('fa_layernorm_bwd', _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dres), _p(dx), _p(dgamma), _p(dbeta),
          rows, C, _stream())
    return dx


def bn_stats(x, B, C, S):
    sums = torch.zeros(C, 2, device=x.device, dtype=torch.float64)
    _call('fa_bn_stats', _p(x), _p(sums), B, C, S, _stream())
    return sums


def bn_apply(x, scale, shift, slope, B, C, S, want_y=False, want_pool=True):
    y = torch.empty_like(x) if want_y else None
    pooled = torch.empty(B, C, device=x.device, dtype=torch.float32) if want_pool else None
    _call('fa_bn_apply', _p(x), _p(scale), _p(shift), slope, _p(y), _p(pooled), B, C, S, _stream())
    return y, pooled


def bn_bwd(x, mean, rstd, scale, shift, slope, dy, dpooled, B, C, S, training=True):
    red = None
    if training:
        red = torch.zeros(C, 2, device=x.device, dtype=torch.float64)
        _call('fa_bn_bwd_reduce', _p(x), _p(mean), _p(rstd), _p(scale), _p(shift), slope, _p(dy), _p(dpooled), _p(red),
              B, C, S, _stream())
    dx = torch.empty_like(x)
    _call('fa_bn_bwd_apply', _p(x), _p(mean), _p(rstd), _p(scale), _p(shift), slope, _p(dy), _p(dpooled), _p(red),
          _p(dx), B, C, S, _stream())
    return dx, red


def bn_tokens_stats(x, T, C):
    sums = torch.zeros(C, 2, device=x.device, dtype=torch.float64)
    _call('fa_bn_tokens_stats', _p(x), _p(sums), T, C, _stream())
    return sums


def bn_tokens_apply(x, scale, shift, res, slope, T, C):
    _f32(x, res)
    y = torch.empty_like(x)
    _call('fa_bn_tokens_apply', _p(x), _p(scale), _p(shift), _p(res), slope, _p(y), T, C, _stream())
    return y


def bn_tokens_bwd(x, mean, rstd, scale, yout, slope, dy, T, C, training=True, want_dres=False):
    _f32(x, dy, yout)
    red = torch.zeros(C, 2, device=x.device, dtype=torch.float64)
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    _call('fa_bn_tokens_bwd', _p(x), _p(mean), _p(rstd), _p(scale), _p(yout), slope, _p(dy), _p(red), _p(dx), _p(dres), T,
          C, int(training), _stream())
    return dx, dres, red


def token_mean_fwd(x):
    _f32(x)
    B, HW, C = x.shape
    out = torch.empty(B, C, device=x.device, dtype=torch.float32)
    _call('fa_token_mean_fwd', _p(x), _p(out), B, HW, C, _stream())
    return out


def token_mean_bwd(dy, B, HW, C):
    _f32(dy)
    dx = torch.empty(B, HW, C, device=dy.device, dtype=torch.float32)
    _call('fa_token_mean_bwd', _p(dy), _p(dx), B, HW, C, _stream())
    return dx


# ----------------------------------------------------------------------------- conv pieces
def dwconv_fwd(h1, w, b, B, H, W, C, want_act=True, u2_mode=0):
    """(u2, h2) = (dwconv(h1) + b, gelu(u2)).  u2_mode=1: the first output is gelu'(u2) (what the backward multiplies by);
    u2_mode=None: u2 is not stored at all (inference)."""
    _f32(h1, w, b)
    u2 = torch.empty_like(h1) if u2_mode is not None else None
    h2 = torch.empty_like(h1) if want_act else None
    if BYTE_COUNTER[0] is not None:            # read h1, write the one or two outputs
        BYTE_COUNTER[0] += 4 * h1.numel() * (1 + (u2 is not None) + (h2 is not None))
    _call('fa_dwconv3x3_fwd', _p(h1), _p(w), _p(b), _p(u2), _p(h2), u2_mode or 0, B, H, W, C, _stream())
    return u2, h2


def dwconv_bwd(du2, h1, u1, w, dw, db, B, H, W, C):
    """h1=None: the kernel recomputes h1 = gelu(u1) for the weight gradient instead of reading it."""
    _f32(du2, h1, u1, w)
    du1 = torch.empty_like(du2)
    if BYTE_COUNTER[0] is not None:            # read du2 (+ h1) (+ u1), write du1
        BYTE_COUNTER[0] += 4 * du2.numel() * (2 + (h1 is not None) + (u1 is not None))
    _call('fa_dwconv3x3_bwd', _p(du2), _p(h1), _p(u1), _p(w), _p(du1), _p(dw), _p(db), B, H, W, C, _stream())
    return du1


def im2col(x, B, H, W, C, kh, kw, stride, pad, nchw_in=False):
    _f32(x)
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    col = torch.empty(B * Ho * Wo, kh * kw * C, device=x.device, dtype=torch.float32)
    _call('fa_im2col', _p(x), _p(col), B, H, W, C, kh, kw, stride, pad, int(nchw_in), _stream())
    return col


def col2im(col, B, H, W, C, kh, kw, stride, pad):
    _f32(col)
    dx = torch.empty(B, H * W, C, device=col.device, dtype=torch.float32)
    _call('fa_col2im', _p(col), _p(dx), B, H, W, C, kh, kw, stride, pad, _stream())
    return dx


def conv3x3_out_fwd(t, wk, bias, ximg, B, H, W, C, Co):
    _f32(t, wk, bias, ximg)
    out = torch.empty(B, Co, H * W, device=t.device, dtype=torch.float32)
    _call('fa_conv3x3_out_fwd', _p(t), _p(wk), _p(bias), _p(ximg), _p(out), B, H, W, C, Co, _stream())
    return out


def conv3x3_out_bwd(t, wk, dout, B, H, W, C, Co, want_dt=True):
    _f32(t, wk, dout)
    dt = torch.empty(B, H * W, C, device=t.device, dtype=torch.float32) if want_dt else None
    dW = torch.zeros(Co, 9 * C, device=t.device, dtype=torch.float32)
    db = torch.zeros(Co, device=t.device, dtype=torch.float32)
    _call('fa_conv3x3_out_bwd', _p(t), _p(wk), _p(dout), _p(dt), _p(dW), _p(db), B, H, W, C, Co, _stream())
    return dt, dW, db


def pixel_shuffle2_fwd(g, y, B, H, W, Co):
    _f32(g)
    _call('fa_pixel_shuffle2_fwd', _p(g), _p(y), y.stride(0), B, H, W, Co, _stream())


def pixel_shuffle2_bwd(dy, dg, B, H, W, Co):
    _f32(dg)
    _call('fa_pixel_shuffle2_bwd', _p(dy), dy.stride(0), _p(dg), B, H, W, Co, _stream())


def copy2d(src, dst):
    rows, cols, lds = _rows2d(src)
    _, _, ldd = _rows2d(dst)
    _call('fa_copy2d', _p(src), lds, _p(dst), ldd, rows, cols, _stream())


def add2d(a, b, dst):
    rows, cols, lda = _rows2d(a)
    _, _, ldb = _rows2d(b)
    _, _, ldd = _rows2d(dst)
    _call('fa_add2d', _p(a), lda, _p(b), ldb, _p(dst), ldd, rows, cols, _stream())


def scale_rows(g, rowscale, rows_per_scale):
    """g * rowscale[row // rows_per_scale]; returns g itself when rowscale is None."""
    if rowscale is None:
        return g
    _f32(g, rowscale)
    out = torch.empty_like(g)
    cols = g.shape[-1]
    _call('fa_scale_rows', _p(g), _p(rowscale), rows_per_scale, _p(out), g.numel() // cols, cols, _stream())
    return out


def tokens_to_nchw(t, res, B, HW, C):
    _f32(t, res)
    y = torch.empty(B, C, HW, device=t.device, dtype=torch.float32)
    _call('fa_tokens_to_nchw', _p(t), _p(res), _p(y), B, HW, C, _stream())
    return y


def nchw_to_tokens(x, B, HW, C):
    _f32(x)
    t = torch.empty(B, HW, C, device=x.device, dtype=torch.float32)
    _call('fa_nchw_to_tokens', _p(x), _p(t), B, HW, C, _stream())
    return t


# ----------------------------------------------------------------------------- DGRN pieces
def dcn_im2col(x, om, B, H, W, C, col=None):
    """om [T, 27 or 32] (row pitch = its width)."""
    _f32(x, om)
    col = torch.empty(B * H * W, 9 * C, device=x.device, dtype=torch.float32) if col is None else col
    _call('fa_dcn_im2col', _p(x), _p(om), om.shape[-1], _p(col), B, H, W, C, _stream())
    return col


def dcn_col2im(x, om, dcol, B, H, W, C):
    _f32(x, om, dcol)
    dx = torch.zeros_like(x)
    dom = torch.empty_like(om) if om.shape[-1] == 27 else torch.zeros_like(om)     # pad columns of a 32-wide dom stay 0
    _call('fa_dcn_col2im', _p(x), _p(om), om.shape[-1], _p(dcol), _p(dx), _p(dom), B, H, W, C, _stream())
    return dx, dom


def sft_fuse_fwd(x, dcn, gamma, beta, slope):
    _f32(x, dcn, gamma, beta)
    out = torch.empty_like(x)
    _call('fa_sft_fuse_fwd', _p(x), _p(dcn), _p(gamma), _p(beta), _p(out), x.numel(), slope, _stream())
    return out


def sft_fuse_bwd(x, dcn, gamma, beta, dout, slope):
    _f32(x, dcn, gamma, beta, dout)
    dx, ddcn, dgamma, dbeta = (torch.empty_like(x) for _ in range(4))
    _call('fa_sft_fuse_bwd', _p(x), _p(dcn), _p(gamma), _p(beta), _p(dout), _p(dx), _p(ddcn), _p(dgamma), _p(dbeta),
          x.numel(), slope, _stream())
    return dx, ddcn, dgamma, dbeta


# ----------------------------------------------------------------------------- elementwise / optimiser
def act_fwd(x, act, p=0.0):
    _f32(x)
    y = torch.empty_like(x)
    _call('fa_act_fwd', _p(x), _p(y), x.numel(), act, p, _stream())
    return y


def act_bwd(dy, x, act, p=0.0):
    _f32(dy, x)
    dx = torch.empty_like(x)
    _call('fa_act_bwd', _p(dy), _p(x), _p(dx), x.numel(), act, p, _stream())
    return dx


def l1_loss(a, b, want_grad=True, gscale=1.0):
    _f32(a, b)
    loss = torch.zeros(1, device=a.device, dtype=torch.float32)
    grad = torch.empty_like(a) if want_grad else None
    _call('fa_l1_loss', _p(a), _p(b), _p(loss), _p(grad), a.numel(), gscale, _stream())
    return loss, grad


def spectral_l1(a, b, bob, nbands, want_grad=True, gscale=1.0):
    """(loss[1], grad) of mean|D(a) - D(b)| over the [nbands, ..., n, n, 2] spectrum stack (train.py:91)."""
    _f32(a, b)
    n = a.shape[-1]
    assert a.shape == b.shape and a.shape[-2] == n
    loss = torch.zeros(1, device=a.device, dtype=torch.float32)
    grad = torch.empty_like(a) if want_grad else None
    _call('fa_spectral_l1', _p(a), _p(b), _p(loss), _p(grad), a.numel() // (n * n), n, _p(bob), nbands, gscale, _stream())
    return loss, grad


def crop_augment(pool, meta, sigma, noise, P):
    """(degraded, clean) [B,3,P,P] training patches from the uint8 image pool (fa_crop_augment)."""
    B = meta.shape[0]
    assert pool.dtype == torch.uint8 and meta.dtype == torch.int64 and meta.shape[1] == 8 and meta.is_contiguous()
    _f32(sigma, noise)
    deg = torch.empty(B, 3, P, P, device=pool.device, dtype=torch.float32)
    clean = torch.empty_like(deg)
    _call('fa_crop_augment', _p(pool), _p(meta), _p(sigma), _p(noise), _p(deg), _p(clean), B, P, _stream())
    return deg, clean


def momentum_update(k, q, m):
    _f32(k, q)
    _call('fa_momentum_update', _p(k), _p(q), k.numel(), m, _stream())


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    _f32(p, g, m, v)
    _call('fa_adam_step', _p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, step, grad_scale, _stream())


def adam_tick(state):
    """state = {lr (fp32), step count (int32)} in device memory: count += 1 (fa_adam_tick; graph-replay safe)."""
    _f32(state)
    _call('fa_adam_tick', _p(state), _stream())


def adam_step_state(p, g, m, v, state, beta1, beta2, eps, grad_scale=1.0):
    """Adam with lr and the step count read from the 2-word device tensor ``state``; the bias corrections are derived
    inside the kernel, so a captured graph of the step never reads host memory."""
    _f32(p, g, m, v, state)
    _call('fa_adam_step_state', _p(p), _p(g), _p(m), _p(v), p.numel(), _p(state), beta1, beta2, eps, grad_scale, _stream())


def round_tf32(src, dst=None):
    """src rounded to the nearest TF32 (a new tensor unless dst is given)."""
    _f32(src)
    dst = torch.empty_like(src) if dst is None else dst
    _call('fa_round_tf32', _p(src), _p(dst), src.numel(), _stream())
    return dst


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        _f32(t)
        arr[i] = t.data_ptr()
    return arr


def band_coef_fwd(stats, params, out, band):
    """out[:, :, band] = lambda predictor of one band; params = (ln_w, ln_b, fc_w, fc_b, w0, b0, w2, b2)."""
    _f32(stats)
    B, D = stats.shape
    heads, nb = out.shape[1], out.shape[2]
    _call('fa_band_coef_fwd', _p(stats), _ptr_array(params), ctypes.c_void_p(out.data_ptr() + 4 * band), B, D, heads,
          heads * nb, nb, _stream())


def band_coef_bwd(stats, params, dout, band, dstats, grads):
    _f32(stats, dout, dstats)
    B, D = stats.shape
    heads, nb = dout.shape[1], dout.shape[2]
    _call('fa_band_coef_bwd', _p(stats), _ptr_array(params), ctypes.c_void_p(dout.data_ptr() + 4 * band), heads * nb, nb,
          _p(dstats), _ptr_array(grads), B, D, heads, _stream())
