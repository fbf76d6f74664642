"""Block-level autograd nodes of the LeWin transformer path.

One ``torch.autograd.Function`` per LeWin block (instead of ~40 fine-grained torch ops in the
reference, encoder_Uformer.py:597-682 / decoder_Uformer.py:618-756): the forward enqueues ~10
libfreqair kernels and keeps the activations the hand-written backward needs; roll /
window_partition / window_reverse never materialise (the attention kernels gather windows from
image-order tokens), DropPath is a per-sample scale folded into the residual GEMM epilogue, and both
GELUs ride GEMM / depthwise-conv epilogues.
"""
import torch

from .. import ops


def _z(t):
    return torch.zeros_like(t)


def linear_grads(g, x, W, dW, db, want_dx=True, dx=None, accumulate_dx=False):
    """Backward of y = x W^T + b for 2-D row-major g [T,N], x [T,K], W [N,K]: fills dW (+=), db, returns dx."""
    if db is not None:
        ops.colsum(g, db)
    ops.gemm(g, x, dW, transA=True, transB=False, accumulate=True)
    if not want_dx:
        return None
    if dx is None:
        dx = torch.empty_like(x)
    ops.gemm(g, W, dx, transB=False, accumulate=accumulate_dx)
    return dx


# ----------------------------------------------------------------------------- LeFF
def leff_fwd(xn2, w1, b1, wdw, bdw, w2, b2, B, H, W, residual, dp_scale, save):
    T, C = xn2.shape
    Ch = w1.shape[0]
    u1 = torch.empty(T, Ch, device=xn2.device, dtype=torch.float32)
    h1 = torch.empty_like(u1)
    ops.gemm(xn2, w1, h1, bias=b1, act=ops.ACT_GELU, preact=u1)
    u2, h2 = ops.dwconv_fwd(h1, wdw, bdw, B, H, W, Ch)
    out = torch.empty(T, w2.shape[0], device=xn2.device, dtype=torch.float32)
    ops.gemm(h2, w2, out, bias=b2, rowscale=dp_scale, rows_per_scale=H * W, residual=residual)
    if save is not None:
        save.update(u1=u1, h1=h1, u2=u2, h2=h2)
    return out


def leff_bwd(gs, sv, xn2, w1, wdw, w2, B, H, W):
    """gs: gradient wrt the LeFF output (DropPath scale already applied). Returns dxn2 and param grads."""
    dW2, db2 = _z(w2), torch.empty(w2.shape[0], device=gs.device)
    ops.colsum(gs, db2)
    ops.gemm(gs, sv['h2'], dW2, transA=True, transB=False, accumulate=True)
    du2 = torch.empty_like(sv['u2'])
    ops.gemm(gs, w2, du2, transB=False, aux=sv['u2'], aux_act=ops.ACT_GELU)
    dwdw, dbdw = _z(wdw), torch.zeros(wdw.shape[0], device=gs.device)
    du1 = ops.dwconv_bwd(du2, sv['h1'], sv['u1'], wdw, dwdw, dbdw, B, H, W, wdw.shape[0])
    dW1, db1 = _z(w1), torch.empty(w1.shape[0], device=gs.device)
    dxn2 = linear_grads(du1, xn2, w1, dW1, db1)
    return dxn2, (dW1, db1, dwdw, dbdw, dW2, db2)


class DecoderBlockFn(torch.autograd.Function):
    """LeWinTransformerBlock.forward of the decoder, plain path (decoder_Uformer.py:618-756) with
    WindowAttention (:235-299).  cfg = (B, H, W, heads, shift, band_of_bin, nbands)."""

    @staticmethod
    def forward(ctx, cfg, x, coef, dp_a, dp_m, n1w, n1b, table, wq, bq, wkv, bkv, wp, bp, n2w, n2b, w1, b1, wdw, bdw,
                w2, b2):
        B, H, W, heads, shift, bob, nbands = cfg
        C = x.shape[-1]
        T = B * H * W
        hd = C // heads
        x2d = x.reshape(T, C)
        xn, mean1, rstd1 = ops.layernorm_fwd(x2d, n1w, n1b)
        qkv = torch.empty(T, 3 * C, device=x.device, dtype=torch.float32)
        ops.gemm(xn, wq, qkv[:, :C], bias=bq)
        ops.gemm(xn, wkv, qkv[:, C:], bias=bkv)
        o = torch.empty(T, C, device=x.device, dtype=torch.float32)
        ops.win_attn_fwd(qkv[:, :C], qkv[:, C:], o, B, H, W, heads, hd, shift, hd ** -0.5, table, coef, heads, bob,
                         nbands)
        x1 = torch.empty(T, C, device=x.device, dtype=torch.float32)
        ops.gemm(o, wp, x1, bias=bp, rowscale=dp_a, rows_per_scale=H * W, residual=x2d)
        xn2, mean2, rstd2 = ops.layernorm_fwd(x1, n2w, n2b)
        sv = {}
        x2 = leff_fwd(xn2, w1, b1, wdw, bdw, w2, b2, B, H, W, x1, dp_m, sv)
        ctx.cfg = cfg
        ctx.has = (coef is not None, dp_a is not None, dp_m is not None)
        ctx.save_for_backward(x2d, mean1, rstd1, xn, qkv, o, x1, mean2, rstd2, xn2, sv['u1'], sv['h1'], sv['u2'],
                              sv['h2'], coef, dp_a, dp_m, n1w, table, wq, wkv, wp, n2w, w1, wdw, w2)
        return x2.view(B, H * W, C)

    @staticmethod
    def backward(ctx, dx2):
        (x2d, mean1, rstd1, xn, qkv, o, x1, mean2, rstd2, xn2, u1, h1, u2, h2, coef, dp_a, dp_m, n1w, table, wq, wkv,
         wp, n2w, w1, wdw, w2) = ctx.saved_tensors
        B, H, W, heads, shift, bob, nbands = ctx.cfg
        T, C = x2d.shape
        hd = C // heads
        g = dx2.reshape(T, C).contiguous()
        gs = ops.scale_rows(g, dp_m, H * W)
        dxn2, (dW1, db1, dwdw, dbdw, dW2, db2) = leff_bwd(gs, dict(u1=u1, h1=h1, u2=u2, h2=h2), xn2, w1, wdw, w2, B, H, W)
        dn2w, dn2b = _z(n2w), _z(n2w)
        g1 = ops.layernorm_bwd(dxn2, x1, mean2, rstd2, n2w, g, dn2w, dn2b)
        gs1 = ops.scale_rows(g1, dp_a, H * W)
        dWp, dbp = _z(wp), torch.empty(C, device=g.device)
        do = linear_grads(gs1, o, wp, dWp, dbp)
        dq = torch.empty(T, C, device=g.device)
        dkv = torch.empty(T, 2 * C, device=g.device)
        dtable = _z(table)
        dcoef = _z(coef) if coef is not None else None
        ops.win_attn_bwd(qkv[:, :C], qkv[:, C:], do, dq, dkv, B, H, W, heads, hd, shift, hd ** -0.5, table, dtable, coef,
                         heads, dcoef, bob, nbands)
        dWq, dbq, dWkv, dbkv = _z(wq), torch.empty(C, device=g.device), _z(wkv), torch.empty(2 * C, device=g.device)
        dxn = linear_grads(dq, xn, wq, dWq, dbq)
        linear_grads(dkv, xn, wkv, dWkv, dbkv, dx=dxn, accumulate_dx=True)
        dn1w, dn1b = _z(n1w), _z(n1w)
        dx = ops.layernorm_bwd(dxn, x2d, mean1, rstd1, n1w, g1, dn1w, dn1b)
        return (None, dx.view(B, H * W, C), dcoef, None, None, dn1w, dn1b, dtable, dWq, dbq, dWkv, dbkv, dWp, dbp, dn2w,
                dn2b, dW1, db1, dwdw, dbdw, dW2, db2)


class EncoderBlockFn(torch.autograd.Function):
    """LeWinTransformerBlock.forward of the encoder (encoder_Uformer.py:597-682): intra- then inter-band
    joint attention (each with its own projections) or the plain 'origin' window attention.
    cfg = (L, B, H, W, heads, shift, msa);  x is [(L*B), HW, C].  For msa='origin' the second attention's
    arguments are None."""

    @staticmethod
    def forward(ctx, cfg, x, dp_a, dp_m, n1w, n1b, tabA, wqA, bqA, wkvA, bkvA, wpA, bpA, tabB, wqB, bqB, wkvB, bkvB, wpB,
                bpB, n2w, n2b, w1, b1, wdw, bdw, w2, b2):
        L, B, H, W, heads, shift, msa = cfg
        C = x.shape[-1]
        LB = x.shape[0]
        T = LB * H * W
        hd = C // heads
        scale = hd ** -0.5
        x2d = x.reshape(T, C)
        xn, mean1, rstd1 = ops.layernorm_fwd(x2d, n1w, n1b)
        qkvA = torch.empty(T, 3 * C, device=x.device, dtype=torch.float32)
        ops.gemm(xn, wqA, qkvA[:, :C], bias=bqA)
        ops.gemm(xn, wkvA, qkvA[:, C:], bias=bkvA)
        oA = torch.empty(T, C, device=x.device, dtype=torch.float32)
        x1 = torch.empty(T, C, device=x.device, dtype=torch.float32)
        if msa == 'origin':
            ops.win_attn_fwd(qkvA[:, :C], qkvA[:, C:], oA, LB, H, W, heads, hd, shift, scale, tabA, None, heads, None, 0)
            ops.gemm(oA, wpA, x1, bias=bpA, rowscale=dp_a, rows_per_scale=H * W, residual=x2d)
            yA = qkvB = oB = None
        else:
            ops.joint_attn_fwd(qkvA[:, :C], qkvA[:, C:], oA, L, B, H, W, heads, hd, shift, scale, tabA, 0)
            yA = torch.empty(T, C, device=x.device, dtype=torch.float32)
            ops.gemm(oA, wpA, yA, bias=bpA)
            qkvB = torch.empty(T, 3 * C, device=x.device, dtype=torch.float32)
            ops.gemm(yA, wqB, qkvB[:, :C], bias=bqB)
            ops.gemm(yA, wkvB, qkvB[:, C:], bias=bkvB)
            oB = torch.empty(T, C, device=x.device, dtype=torch.float32)
            ops.joint_attn_fwd(qkvB[:, :C], qkvB[:, C:], oB, L, B, H, W, heads, hd, shift, scale, tabB, 1)
            ops.gemm(oB, wpB, x1, bias=bpB, rowscale=dp_a, rows_per_scale=H * W, residual=x2d)
        xn2, mean2, rstd2 = ops.layernorm_fwd(x1, n2w, n2b)
        sv = {}
        x2 = leff_fwd(xn2, w1, b1, wdw, bdw, w2, b2, LB, H, W, x1, dp_m, sv)
        ctx.cfg = cfg
        ctx.save_for_backward(x2d, mean1, rstd1, xn, qkvA, oA, yA, qkvB, oB, x1, mean2, rstd2, xn2, sv['u1'], sv['h1'],
                              sv['u2'], sv['h2'], dp_a, dp_m, n1w, tabA, wqA, wkvA, wpA, tabB, wqB, wkvB, wpB, n2w, w1,
                              wdw, w2)
        return x2.view(LB, H * W, C)

    @staticmethod
    def backward(ctx, dx2):
        (x2d, mean1, rstd1, xn, qkvA, oA, yA, qkvB, oB, x1, mean2, rstd2, xn2, u1, h1, u2, h2, dp_a, dp_m, n1w, tabA,
         wqA, wkvA, wpA, tabB, wqB, wkvB, wpB, n2w, w1, wdw, w2) = ctx.saved_tensors
        L, B, H, W, heads, shift, msa = ctx.cfg
        T, C = x2d.shape
        LB = T // (H * W)
        hd = C // heads
        scale = hd ** -0.5
        dev = x2d.device
        g = dx2.reshape(T, C).contiguous()
        gs = ops.scale_rows(g, dp_m, H * W)
        dxn2, (dW1, db1, dwdw, dbdw, dW2, db2) = leff_bwd(gs, dict(u1=u1, h1=h1, u2=u2, h2=h2), xn2, w1, wdw, w2, LB, H, W)
        dn2w, dn2b = _z(n2w), _z(n2w)
        g1 = ops.layernorm_bwd(dxn2, x1, mean2, rstd2, n2w, g, dn2w, dn2b)
        gs1 = ops.scale_rows(g1, dp_a, H * W)
        dq = torch.empty(T, C, device=dev)
        dkv = torch.empty(T, 2 * C, device=dev)
        gB = [None] * 7
        if msa == 'origin':
            dWpA, dbpA = _z(wpA), torch.empty(C, device=dev)
            doA = linear_grads(gs1, oA, wpA, dWpA, dbpA)
            dtabA = _z(tabA)
            ops.win_attn_bwd(qkvA[:, :C], qkvA[:, C:], doA, dq, dkv, LB, H, W, heads, hd, shift, scale, tabA, dtabA, None,
                             heads, None, None, 0)
        else:
            dWpB, dbpB = _z(wpB), torch.empty(C, device=dev)
            doB = linear_grads(gs1, oB, wpB, dWpB, dbpB)
            dtabB = _z(tabB)
            ops.joint_attn_bwd(qkvB[:, :C], qkvB[:, C:], doB, dq, dkv, L, B, H, W, heads, hd, shift, scale, tabB, dtabB, 1)
            dWqB, dbqB, dWkvB, dbkvB = _z(wqB), torch.empty(C, device=dev), _z(wkvB), torch.empty(2 * C, device=dev)
            dyA = linear_grads(dq, yA, wqB, dWqB, dbqB)
            linear_grads(dkv, yA, wkvB, dWkvB, dbkvB, dx=dyA, accumulate_dx=True)
            gB = [dtabB, dWqB, dbqB, dWkvB, dbkvB, dWpB, dbpB]
            dWpA, dbpA = _z(wpA), torch.empty(C, device=dev)
            doA = linear_grads(dyA, oA, wpA, dWpA, dbpA)
            dtabA = _z(tabA)
            dq = torch.empty(T, C, device=dev)
            dkv = torch.empty(T, 2 * C, device=dev)
            ops.joint_attn_bwd(qkvA[:, :C], qkvA[:, C:], doA, dq, dkv, L, B, H, W, heads, hd, shift, scale, tabA, dtabA, 0)
        dWqA, dbqA, dWkvA, dbkvA = _z(wqA), torch.empty(C, device=dev), _z(wkvA), torch.empty(2 * C, device=dev)
        dxn = linear_grads(dq, xn, wqA, dWqA, dbqA)
        linear_grads(dkv, xn, wkvA, dWkvA, dbkvA, dx=dxn, accumulate_dx=True)
        dn1w, dn1b = _z(n1w), _z(n1w)
        dx = ops.layernorm_bwd(dxn, x2d, mean1, rstd1, n1w, g1, dn1w, dn1b)
        return (None, dx.view(LB, H * W, C), None, None, dn1w, dn1b, dtabA, dWqA, dbqA, dWkvA, dbkvA, dWpA, dbpA, *gB,
                dn2w, dn2b, dW1, db1, dwdw, dbdw, dW2, db2)


# ----------------------------------------------------------------------------- dense layers as autograd nodes
class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) (+ residual) on the last dim (nn.Linear + optional LeakyReLU/GELU epilogue)."""

    @staticmethod
    def forward(ctx, x, W, b, act, act_param, residual):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        y = torch.empty(x2.shape[0], W.shape[0], device=x.device, dtype=torch.float32)
        pre = torch.empty_like(y) if (act == ops.ACT_GELU or (act != ops.ACT_NONE and residual is not None)) else None
        r2 = residual.reshape(-1, W.shape[0]).contiguous() if residual is not None else None
        ops.gemm(x2, W, y, bias=b, act=act, act_param=act_param, preact=pre, residual=r2)
        ctx.act = (act, act_param)
        ctx.has_bias = b is not None
        ctx.has_res = residual is not None
        ctx.save_for_backward(x2, W, pre if pre is not None else (y if act != ops.ACT_NONE else None))
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], W.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, W, pre = ctx.saved_tensors
        act, ap = ctx.act
        g = dy.reshape(-1, W.shape[0]).contiguous()
        dres = dy if ctx.has_res else None
        if act != ops.ACT_NONE:
            # LeakyReLU: sign(out) == sign(pre-activation), so the saved output serves as aux
            g = ops.act_bwd(g, pre, act, ap)
        dW = _z(W)
        db = torch.empty(W.shape[0], device=g.device) if ctx.has_bias else None
        dx = linear_grads(g, x2, W, dW, db, want_dx=ctx.needs_input_grad[0])
        return (dx.view(ctx.xshape) if dx is not None else None), dW, db, None, None, dres


def linear(x, W, b=None, act=ops.ACT_NONE, act_param=0.0, residual=None):
    return LinearFn.apply(x, W, b, act, act_param, residual)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        xc = x.contiguous()
        y, mean, rstd = ops.layernorm_fwd(xc, w, b)
        ctx.save_for_backward(xc, mean, rstd, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, w = ctx.saved_tensors
        dw, db = _z(w), _z(w)
        dx = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, w, None, dw, db)
        return dx, dw, db


def layer_norm(x, w, b):
    return LayerNormFn.apply(x, w, b)
