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


# Direct gradient sinks.  trainer.TrainStep makes every parameter's .grad a view of one flat, per-step-zeroed buffer and
# marks the parameter (``p._fa_direct = True``); the block-level backward then ACCUMULATES weight / bias / norm gradients
# straight into those views (GEMM epilogue accumulate, atomics) and hands autograd ``None``, which removes ~2 000
# AccumulateGrad add kernels and as many zero-filled temporaries per step.  ``p._fa_ready(p)`` (set by TrainStep when it
# runs data-parallel) tells the bucketed all-reduce that a gradient is complete: the post-accumulate hooks no longer fire
# for these parameters.  Both switches live on the PARAMETER, not in this module, so two TrainSteps - or a TrainStep
# beside a plain optim.Adam loop on another net - do not rewire each other.  A parameter without the mark gets its
# gradient through autograd as usual (what the parity tests exercise).  DIRECT_GRAD = False switches the sinks off
# process-wide (tests only: direct vs autograd-routed gradients).
DIRECT_GRAD = True
FOLD_DROPPATH = True          # fold the DropPath backward scale into the consuming contractions (see _fold)


def _sink(p):
    if (DIRECT_GRAD and p is not None and getattr(p, '_fa_direct', False) and p.requires_grad and p.grad is not None
            and p.grad.is_contiguous()):
        return p.grad
    return None


def _wbuf(p):
    """(buffer to accumulate a weight-like gradient into, value to return to autograd)"""
    sk = _sink(p)
    if sk is not None:
        return sk, None
    z = torch.zeros_like(p)
    return z, z


def _bias_grad(g, p):
    """Column sums of g into the bias gradient; returns the value for autograd (None when written in place)."""
    if p is None:
        return None
    sk = _sink(p)
    if sk is not None:
        ops.colsum(g, sk, accumulate=True)
        return None
    db = torch.empty_like(p)
    ops.colsum(g, db)
    return db


def _ready(*ps):
    for p in ps:
        if p is not None:
            cb = getattr(p, '_fa_ready', None)
            if cb is not None and _sink(p) is not None:
                cb(p)


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


class _QKV:
    """q [T, C] and kv [T, 2C] as two dense blocks of one allocation.  (As column slices of a [T, 3C] buffer their rows
    start at 112-byte offsets at the C = 28 levels and the GEMMs' stores split 32-byte sectors: measured 2x on those
    projections.)"""

    def __init__(self, T=None, C=None, device=None, buf=None):
        self.buf = torch.empty(T * 3 * C, device=device, dtype=torch.float32) if buf is None else buf

    def views(self, T, C):
        return self.buf[:T * C].view(T, C), self.buf[T * C:].view(T, 2 * C)


def _fold(g, dp, rps):
    """DropPath backward g * dp[sample]: either folded into the consumers (returns g, dp) - the weight-gradient
    contraction scales its A operand per k-block (FaGemmEpilogue.a_kscale), the data-gradient contraction its output rows
    - or, when a sample is not a whole number of 32-token k-blocks, materialised with fa_scale_rows (returns g*dp, None)."""
    if dp is None:
        return g, None
    if FOLD_DROPPATH and rps % 32 == 0:
        return g, dp
    return ops.scale_rows(g, dp, rps), None


def linear_param_grads(g, x, W, b, want_dx=True, dx=None, accumulate_dx=False, dp=None, rps=1, backend=None):
    """linear_grads with the parameter gradients routed through the sinks: returns (dx, dW_ret, db_ret).
    dp / rps: per-sample scale of g (DropPath backward) folded into both contractions.
    backend: fa_gemm backend of both contractions (the LeFF class passes ops.LEFF_BACKEND)."""
    dW, dW_ret = _wbuf(W)
    db, db_ret = _wbuf(b) if b is not None else (None, None)
    # dW += g^T x, and the bias gradient (column sums of g) taken from the same pass over g (FaGemmEpilogue.a_rowsum)
    ops.gemm(g, x, dW, transA=True, transB=False, accumulate=True, a_rowsum=db, a_kscale=dp, a_k_rows_per_scale=rps,
             backend=backend)
    _ready(W, b)
    if not want_dx:
        return None, dW_ret, db_ret
    if dx is None:
        dx = torch.empty_like(x)
    Wr, bx = rn_weight(W, backend)
    ops.gemm(g, Wr, dx, transB=False, accumulate=accumulate_dx, rowscale=dp, rows_per_scale=rps, backend=backend, b_is_tf32=bx)
    return dx, dW_ret, db_ret


# ----------------------------------------------------------------------------- LeFF
# Rounded weight copies.  The LeFF contractions run 1xTF32 (ops.LEFF_BACKEND) and want their weight operand already
# rounded to TF32 (the kernel then skips its in-place rounding pass over the B tile).  trainer.TrainStep keeps ONE rounded
# copy of each flat parameter buffer, refreshed at the start of every step, and marks the parameters with views of it
# (``p._fa_rn``); the views are only trusted while that TrainStep's step is running (RN_OWNER), because nothing else
# (load_state_dict, an external optimiser) keeps them in sync.  Anywhere else the copy is made on the fly.
RN_OWNER = [None]


def rn_weight(W, backend=None):
    """(weight operand for a contraction on fa_gemm backend `backend`, b_is_tf32)"""
    if backend not in (ops.GEMM_1X, ops.GEMM_2X):
        return W, False
    r = getattr(W, '_fa_rn', None)
    if r is not None and RN_OWNER[0] is not None and getattr(W, '_fa_rn_owner', None) == RN_OWNER[0]:
        return r, True
    return ops.round_tf32(W.detach().contiguous()), True


def leff_fwd(xn2, w1, b1, wdw, bdw, w2, b2, B, H, W, residual, dp_scale, save, be=0):
    T, C = xn2.shape
    Ch = w1.shape[0]
    h1 = torch.empty(T, Ch, device=xn2.device, dtype=torch.float32)
    u1 = torch.empty_like(h1) if save is not None else None          # pre-activation: only the backward reads it
    w1r, x1 = rn_weight(w1, be)
    w2r, x2 = rn_weight(w2, be)
    ops.gemm(xn2, w1r, h1, bias=b1, act=ops.ACT_GELU, preact=u1, backend=be, b_is_tf32=x1)
    # saved for the backward: gelu'(u2), not u2 - it only ever multiplies dh2 (leff_bwd); nothing in inference
    u2, h2 = ops.dwconv_fwd(h1, wdw, bdw, B, H, W, Ch, u2_mode=1 if save is not None else None)
    out = torch.empty(T, w2.shape[0], device=xn2.device, dtype=torch.float32)
    ops.gemm(h2, w2r, out, bias=b2, rowscale=dp_scale, rows_per_scale=H * W, residual=residual, backend=be, b_is_tf32=x2)
    if save is not None:
        save.update(u1=u1, u2=u2, h2=h2)          # h1 = gelu(u1) is recomputed by the depthwise-conv backward
    return out


def leff_bwd(gs, sv, xn2, w1, b1, wdw, bdw, w2, b2, B, H, W, dp=None, be=0):
    """gs: gradient wrt the LeFF output; dp: its per-sample DropPath scale when it is folded into the two contractions
    that consume gs (None: already applied). Returns dxn2 and the parameter gradients (None where they were
    accumulated straight into the parameters' .grad buffers)."""
    be = ops.BWD_BACKEND if be == 0 else be        # backward contractions: ops.BWD_BACKEND unless the forward class is reduced already
    dW2, dW2r = _wbuf(w2)
    db2b, db2 = _wbuf(b2)
    ops.gemm(gs, sv['h2'], dW2, transA=True, transB=False, accumulate=True, a_rowsum=db2b, a_kscale=dp,
             a_k_rows_per_scale=H * W, backend=be)
    du2 = torch.empty_like(sv['u2'])
    w2r, x2 = rn_weight(w2, be)
    ops.gemm(gs, w2r, du2, transB=False, aux=sv['u2'], aux_act=ops.ACT_MUL, rowscale=dp, rows_per_scale=H * W,
             backend=be, b_is_tf32=x2)                                                                          # u2 = gelu'
    dwdw, dwdwr = _wbuf(wdw)
    dbdw, dbdwr = _wbuf(bdw)
    du1 = ops.dwconv_bwd(du2, None, sv['u1'], wdw, dwdw, dbdw, B, H, W, wdw.shape[0])      # h1 = gelu(u1) recomputed
    _ready(w2, b2, wdw, bdw)
    dxn2, dW1r, db1r = linear_param_grads(du1, xn2, w1, b1, backend=be)
    return dxn2, (dW1r, db1r, dwdwr, dbdwr, dW2r, db2)


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
        qkv = _QKV(T, C, x.device).buf
        q_, kv_ = _QKV(buf=qkv).views(T, C)
        ops.gemm(xn, wq, q_, bias=bq)
        ops.gemm(xn, wkv, kv_, bias=bkv)
        o = torch.empty(T, C, device=x.device, dtype=torch.float32)
        ops.win_attn_fwd(q_, kv_, o, B, H, W, heads, hd, shift, hd ** -0.5, table, coef, heads, bob,
                         nbands)
        x1 = torch.empty(T, C, device=x.device, dtype=torch.float32)
        ops.gemm(o, wp, x1, bias=bp, rowscale=dp_a, rows_per_scale=H * W, residual=x2d)
        xn2, mean2, rstd2 = ops.layernorm_fwd(x1, n2w, n2b)
        need_bwd = any(ctx.needs_input_grad)
        sv = {} if need_bwd else None                 # inference: the LeFF intermediates u1, u2 are never stored
        x2 = leff_fwd(xn2, w1, b1, wdw, bdw, w2, b2, B, H, W, x1, dp_m, sv, ops.LEFF_BACKEND)
        if not need_bwd:
            return x2.view(B, H * W, C)
        ctx.cfg = cfg
        ctx.has = (coef is not None, dp_a is not None, dp_m is not None)
        ctx.params = (n1w, n1b, table, wq, bq, wkv, bkv, wp, bp, n2w, n2b, w1, b1, wdw, bdw, w2, b2)
        ctx.save_for_backward(x2d, mean1, rstd1, xn, qkv, o, x1, mean2, rstd2, xn2, sv['u1'], sv['u2'],
                              sv['h2'], coef, dp_a, dp_m, n1w, table, wq, wkv, wp, n2w, w1, wdw, w2)
        return x2.view(B, H * W, C)

    @staticmethod
    def backward(ctx, dx2):
        (x2d, mean1, rstd1, xn, qkv, o, x1, mean2, rstd2, xn2, u1, u2, h2, coef, dp_a, dp_m, n1w, table, wq, wkv,
         wp, n2w, w1, wdw, w2) = ctx.saved_tensors
        (P_n1w, P_n1b, P_table, P_wq, P_bq, P_wkv, P_bkv, P_wp, P_bp, P_n2w, P_n2b, P_w1, P_b1, P_wdw, P_bdw, P_w2,
         P_b2) = ctx.params
        B, H, W, heads, shift, bob, nbands = ctx.cfg
        T, C = x2d.shape
        hd = C // heads
        g = dx2.reshape(T, C).contiguous()
        gs, dpf = _fold(g, dp_m, H * W)
        dxn2, (dW1, db1, dwdw, dbdw, dW2, db2) = leff_bwd(gs, dict(u1=u1, u2=u2, h2=h2), xn2, P_w1, P_b1, P_wdw,
                                                         P_bdw, P_w2, P_b2, B, H, W, dp=dpf, be=ops.LEFF_BACKEND)
        dn2w, dn2wr = _wbuf(P_n2w)
        dn2b, dn2br = _wbuf(P_n2b)
        g1 = ops.layernorm_bwd(dxn2, x1, mean2, rstd2, n2w, g, dn2w, dn2b)
        _ready(P_n2w, P_n2b)
        gs1, dpf = _fold(g1, dp_a, H * W)
        do, dWp, dbp = linear_param_grads(gs1, o, P_wp, P_bp, dp=dpf, rps=H * W, backend=ops.BWD_BACKEND)
        dq = torch.empty(T, C, device=g.device)
        dkv = torch.empty(T, 2 * C, device=g.device)
        dtable, dtabler = _wbuf(P_table)
        dcoef = _z(coef) if coef is not None else None
        q_, kv_ = _QKV(buf=qkv).views(T, C)
        ops.win_attn_bwd(q_, kv_, do, dq, dkv, B, H, W, heads, hd, shift, hd ** -0.5, table, dtable, coef,
                         heads, dcoef, bob, nbands)
        _ready(P_table)
        dxn, dWq, dbq = linear_param_grads(dq, xn, P_wq, P_bq, backend=ops.BWD_BACKEND)
        _, dWkv, dbkv = linear_param_grads(dkv, xn, P_wkv, P_bkv, dx=dxn, accumulate_dx=True, backend=ops.BWD_BACKEND)
        dn1w, dn1wr = _wbuf(P_n1w)
        dn1b, dn1br = _wbuf(P_n1b)
        dx = ops.layernorm_bwd(dxn, x2d, mean1, rstd1, n1w, g1, dn1w, dn1b)
        _ready(P_n1w, P_n1b)
        return (None, dx.view(B, H * W, C), dcoef, None, None, dn1wr, dn1br, dtabler, dWq, dbq, dWkv, dbkv, dWp, dbp,
                dn2wr, dn2br, dW1, db1, dwdw, dbdw, dW2, db2)


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
        qkvA = _QKV(T, C, x.device).buf
        qA_, kvA_ = _QKV(buf=qkvA).views(T, C)
        ops.gemm(xn, wqA, qA_, bias=bqA)
        ops.gemm(xn, wkvA, kvA_, bias=bkvA)
        oA = torch.empty(T, C, device=x.device, dtype=torch.float32)
        x1 = torch.empty(T, C, device=x.device, dtype=torch.float32)
        if msa == 'origin':
            ops.win_attn_fwd(qA_, kvA_, oA, LB, H, W, heads, hd, shift, scale, tabA, None, heads, None, 0)
            ops.gemm(oA, wpA, x1, bias=bpA, rowscale=dp_a, rows_per_scale=H * W, residual=x2d)
            yA = qkvB = oB = None
        else:
            ops.joint_attn_fwd(qA_, kvA_, oA, L, B, H, W, heads, hd, shift, scale, tabA, 0)
            yA = torch.empty(T, C, device=x.device, dtype=torch.float32)
            ops.gemm(oA, wpA, yA, bias=bpA)
            qkvB = _QKV(T, C, x.device).buf
            qB_, kvB_ = _QKV(buf=qkvB).views(T, C)
            ops.gemm(yA, wqB, qB_, bias=bqB)
            ops.gemm(yA, wkvB, kvB_, bias=bkvB)
            oB = torch.empty(T, C, device=x.device, dtype=torch.float32)
            ops.joint_attn_fwd(qB_, kvB_, oB, L, B, H, W, heads, hd, shift, scale, tabB, 1)
            ops.gemm(oB, wpB, x1, bias=bpB, rowscale=dp_a, rows_per_scale=H * W, residual=x2d)
        xn2, mean2, rstd2 = ops.layernorm_fwd(x1, n2w, n2b)
        need_bwd = any(ctx.needs_input_grad)
        sv = {} if need_bwd else None
        x2 = leff_fwd(xn2, w1, b1, wdw, bdw, w2, b2, LB, H, W, x1, dp_m, sv, ops.LEFF_ENC_BACKEND)
        if not need_bwd:
            return x2.view(LB, H * W, C)
        ctx.cfg = cfg
        ctx.params = (n1w, n1b, wqA, bqA, wkvA, bkvA, wpA, bpA, wqB, bqB, wkvB, bkvB, wpB, bpB, n2w, n2b, w1, b1, wdw,
                      bdw, w2, b2)
        ctx.save_for_backward(x2d, mean1, rstd1, xn, qkvA, oA, yA, qkvB, oB, x1, mean2, rstd2, xn2, sv['u1'],
                              sv['u2'], sv['h2'], dp_a, dp_m, n1w, tabA, wqA, wkvA, wpA, tabB, wqB, wkvB, wpB, n2w, w1,
                              wdw, w2)
        return x2.view(LB, H * W, C)

    @staticmethod
    def backward(ctx, dx2):
        (x2d, mean1, rstd1, xn, qkvA, oA, yA, qkvB, oB, x1, mean2, rstd2, xn2, u1, u2, h2, dp_a, dp_m, n1w, tabA,
         wqA, wkvA, wpA, tabB, wqB, wkvB, wpB, n2w, w1, wdw, w2) = ctx.saved_tensors
        (P_n1w, P_n1b, P_wqA, P_bqA, P_wkvA, P_bkvA, P_wpA, P_bpA, P_wqB, P_bqB, P_wkvB, P_bkvB, P_wpB, P_bpB, P_n2w,
         P_n2b, P_w1, P_b1, P_wdw, P_bdw, P_w2, P_b2) = ctx.params
        L, B, H, W, heads, shift, msa = ctx.cfg
        T, C = x2d.shape
        LB = T // (H * W)
        hd = C // heads
        scale = hd ** -0.5
        dev = x2d.device
        g = dx2.reshape(T, C).contiguous()
        gs, dpf = _fold(g, dp_m, H * W)
        dxn2, (dW1, db1, dwdw, dbdw, dW2, db2) = leff_bwd(gs, dict(u1=u1, u2=u2, h2=h2), xn2, P_w1, P_b1, P_wdw,
                                                         P_bdw, P_w2, P_b2, LB, H, W, dp=dpf, be=ops.LEFF_ENC_BACKEND)
        dn2w, dn2wr = _wbuf(P_n2w)
        dn2b, dn2br = _wbuf(P_n2b)
        g1 = ops.layernorm_bwd(dxn2, x1, mean2, rstd2, n2w, g, dn2w, dn2b)
        _ready(P_n2w, P_n2b)
        gs1, dpf = _fold(g1, dp_a, H * W)
        dq = torch.empty(T, C, device=dev)
        dkv = torch.empty(T, 2 * C, device=dev)
        gB = [None] * 7
        if msa == 'origin':
            doA, dWpA, dbpA = linear_param_grads(gs1, oA, P_wpA, P_bpA, dp=dpf, rps=H * W, backend=ops.BWD_BACKEND)
            dtabA = _z(tabA)
            qA_, kvA_ = _QKV(buf=qkvA).views(T, C)
            ops.win_attn_bwd(qA_, kvA_, doA, dq, dkv, LB, H, W, heads, hd, shift, scale, tabA, dtabA, None,
                             heads, None, None, 0)
        else:
            doB, dWpB, dbpB = linear_param_grads(gs1, oB, P_wpB, P_bpB, dp=dpf, rps=H * W, backend=ops.BWD_BACKEND)
            dtabB = _z(tabB)
            qB_, kvB_ = _QKV(buf=qkvB).views(T, C)
            ops.joint_attn_bwd(qB_, kvB_, doB, dq, dkv, L, B, H, W, heads, hd, shift, scale, tabB, dtabB, 1)
            dyA, dWqB, dbqB = linear_param_grads(dq, yA, P_wqB, P_bqB, backend=ops.BWD_BACKEND)
            _, dWkvB, dbkvB = linear_param_grads(dkv, yA, P_wkvB, P_bkvB, dx=dyA, accumulate_dx=True, backend=ops.BWD_BACKEND)
            gB = [dtabB, dWqB, dbqB, dWkvB, dbkvB, dWpB, dbpB]
            doA, dWpA, dbpA = linear_param_grads(dyA, oA, P_wpA, P_bpA, backend=ops.BWD_BACKEND)
            dtabA = _z(tabA)
            dq = torch.empty(T, C, device=dev)
            dkv = torch.empty(T, 2 * C, device=dev)
            qA_, kvA_ = _QKV(buf=qkvA).views(T, C)
            ops.joint_attn_bwd(qA_, kvA_, doA, dq, dkv, L, B, H, W, heads, hd, shift, scale, tabA, dtabA, 0)
        dxn, dWqA, dbqA = linear_param_grads(dq, xn, P_wqA, P_bqA, backend=ops.BWD_BACKEND)
        _, dWkvA, dbkvA = linear_param_grads(dkv, xn, P_wkvA, P_bkvA, dx=dxn, accumulate_dx=True, backend=ops.BWD_BACKEND)
        dn1w, dn1wr = _wbuf(P_n1w)
        dn1b, dn1br = _wbuf(P_n1b)
        dx = ops.layernorm_bwd(dxn, x2d, mean1, rstd1, n1w, g1, dn1w, dn1b)
        _ready(P_n1w, P_n1b)
        return (None, dx.view(LB, H * W, C), None, None, dn1wr, dn1br, dtabA, dWqA, dbqA, dWkvA, dbkvA, dWpA, dbpA, *gB,
                dn2wr, dn2br, dW1, db1, dwdw, dbdw, dW2, db2)


# ----------------------------------------------------------------------------- dense layers as autograd nodes
class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) (+ residual) on the last dim (nn.Linear + optional LeakyReLU/GELU epilogue)."""

    @staticmethod
    def forward(ctx, x, W, b, act, act_param, residual, bwd_backend=0):
        ctx.bwd_backend = bwd_backend
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        y = torch.empty(x2.shape[0], W.shape[0], device=x.device, dtype=torch.float32)
        pre = torch.empty_like(y) if (act == ops.ACT_GELU or (act != ops.ACT_NONE and residual is not None)) else None
        r2 = residual.reshape(-1, W.shape[0]).contiguous() if residual is not None else None
        ops.gemm(x2, W, y, bias=b, act=act, act_param=act_param, preact=pre, residual=r2)
        ctx.act = (act, act_param)
        ctx.has_bias = b is not None
        ctx.has_res = residual is not None
        ctx.params = (W, b)
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
        P_W, P_b = ctx.params
        dx, dW, db = linear_param_grads(g, x2, P_W, P_b, want_dx=ctx.needs_input_grad[0], backend=ctx.bwd_backend)
        return (dx.view(ctx.xshape) if dx is not None else None), dW, db, None, None, dres, None


def linear(x, W, b=None, act=ops.ACT_NONE, act_param=0.0, residual=None, bwd_backend=0):
    """bwd_backend: fa_gemm backend of the two backward contractions (0 = 3xTF32; the Uformer heads pass ops.BWD_BACKEND)."""
    return LinearFn.apply(x, W, b, act, act_param, residual, bwd_backend)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        xc = x.contiguous()
        y, mean, rstd = ops.layernorm_fwd(xc, w, b)
        ctx.save_for_backward(xc, mean, rstd, w)
        ctx.params = (w, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, w = ctx.saved_tensors
        P_w, P_b = ctx.params
        dw, dwr = _wbuf(P_w)
        db, dbr = _wbuf(P_b)
        dx = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, w, None, dw, db)
        _ready(P_w, P_b)
        return dx, dwr, dbr


def layer_norm(x, w, b):
    return LayerNormFn.apply(x, w, b)
