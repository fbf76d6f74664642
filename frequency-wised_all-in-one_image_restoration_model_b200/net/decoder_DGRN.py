"""DGRN - drop-in for the reference's net/decoder_DGRN.py (``DGRN(opt)``, ``forward(x, inter) -> restored``,
identical state_dict keys): head conv -> 5 DGG x 5 DGB x 2 DGM (= 50 DCNv2 + 50 SFT) -> tail conv.

One autograd node per DGM (decoder_DGRN.py:22-32): offset/mask conv, DCNv2 gather + contraction, the two
SFT 1x1-conv MLPs, ``x + dcn + x*gamma + beta`` and the LeakyReLU that DGB applies next (:79,81) fused.
The degradation map's half of every offset conv input (``cat([x, inter])``, deform_conv.py:57-59) is the same
for all 50 DGMs, so its 3x3 patch matrix is gathered once per forward and each DGM contracts it with its own
weight half.  Everything runs on NHWC tokens; NCHW only at the module boundary.
"""
import torch
import torch.nn as nn

from .. import ops
from .convs import Im2colFn, NchwToTokensFn, TokensToNchwFn, conv_tokens, conv_weight_matrix
from .lewin import _z
from .utils.deform_conv import DCN_layer


def default_conv(in_channels, out_channels, kernel_size, bias=True):
    return nn.Conv2d(in_channels, out_channels, kernel_size, padding=(kernel_size // 2), bias=bias)


def _pad_rows(w, rows=32):
    return torch.cat([w, w.new_zeros(rows - w.shape[0], *w.shape[1:])], 0).contiguous()


class DGMFn(torch.autograd.Function):
    """out = lrelu_slope(x + DCN(x, inter) + SFT(x, inter)); slope 1.0 = the bare DGM.

    Offset / mask convolution (Conv2d(2C -> 27, 3x3) on cat([x, inter]), deform_conv.py:57-59): where the geometry
    allows (every DGM of the 128 x 128, 64-channel DGRN) it runs as two IMPLICIT-GEMM convolutions - one over x, one over
    the degradation map - accumulating into a 32-wide om buffer (27 channels + 5 zero pad so that rows are 128 bytes and
    Cout % 4 == 0); its backward takes dom [T, 32] as a 32-channel token tensor: weight gradients by fa_conv3x3_wgrad,
    data gradients by the same implicit kernel with the flipped weights.  No 9x patch matrix of x or of the degradation
    map, and none of their gradients, ever exists.  Small / odd geometries (tests at n_feats 8) keep the explicit
    fa_im2col path with a patch matrix of the degradation map shared by all DGMs (``col_i``)."""

    @staticmethod
    def forward(ctx, x, it, col_i, wom_x, wom_i, bom, wdcn, wg0, wg2, wb0, wb2, H, W, slope):
        B, HW, C = x.shape
        T = B * HW
        x2 = x.reshape(T, C).contiguous()
        it2 = it.reshape(T, -1).contiguous()
        Ci = it2.shape[1]
        implicit = col_i is None
        if implicit:
            wx32, wi32, b32 = _pad_rows(wom_x), _pad_rows(wom_i), _pad_rows(bom)
            om = torch.empty(T, 32, device=x.device, dtype=torch.float32)
            ops.conv3x3_gemm(it2, wi32, om, B, H, W, bias=b32)
            ops.conv3x3_gemm(x2, wx32, om, B, H, W, accumulate=True)
        else:
            colx = ops.im2col(x2, B, H, W, C, 3, 3, 1, 1)
            om = torch.empty(T, 27, device=x.device, dtype=torch.float32)
            ops.gemm(colx, wom_x, om, bias=bom)
            ops.gemm(col_i, wom_i, om, accumulate=True)
        dcol = ops.dcn_im2col(x2, om, B, H, W, C)
        Co = wdcn.shape[0]
        dcn = torch.empty(T, Co, device=x.device, dtype=torch.float32)
        ops.gemm(dcol, wdcn, dcn)
        g1 = torch.empty(T, Co, device=x.device, dtype=torch.float32)
        gamma = torch.empty_like(g1)
        b1 = torch.empty_like(g1)
        beta = torch.empty_like(g1)
        ops.gemm(it2, wg0, g1, act=ops.ACT_LRELU, act_param=0.1)
        ops.gemm(g1, wg2, gamma)
        ops.gemm(it2, wb0, b1, act=ops.ACT_LRELU, act_param=0.1)
        ops.gemm(b1, wb2, beta)
        out = ops.sft_fuse_fwd(x2, dcn, gamma, beta, slope)
        ctx.geom = (B, H, W, C, Ci, slope, implicit)
        ctx.save_for_backward(x2, it2, col_i, om, dcn, g1, gamma, b1, beta, wom_x, wom_i, wdcn, wg0, wg2, wb0, wb2)
        return out.view(B, HW, Co)

    @staticmethod
    def backward(ctx, dout):
        (x2, it2, col_i, om, dcn, g1, gamma, b1, beta, wom_x, wom_i, wdcn, wg0, wg2, wb0, wb2) = ctx.saved_tensors
        B, H, W, C, Ci, slope, implicit = ctx.geom
        T = x2.shape[0]
        dev = x2.device
        dxa, ddcn, dgamma, dbeta = ops.sft_fuse_bwd(x2, dcn, gamma, beta, dout.reshape(T, -1).contiguous(), slope)
        # SFT MLPs
        dit = torch.empty_like(it2)

        def mlp_bwd(dy, h, w0, w2, first):
            dw2 = _z(w2)
            ops.gemm(dy, h, dw2, transA=True, transB=False, accumulate=True)
            dh = torch.empty_like(h)
            ops.gemm(dy, w2, dh, transB=False, aux=h, aux_act=ops.ACT_LRELU, aux_param=0.1)
            dw0 = _z(w0)
            ops.gemm(dh, it2, dw0, transA=True, transB=False, accumulate=True)
            ops.gemm(dh, w0, dit, transB=False, accumulate=not first)
            return dw0, dw2
        dwg0, dwg2 = mlp_bwd(dgamma, g1, wg0, wg2, True)
        dwb0, dwb2 = mlp_bwd(dbeta, b1, wb0, wb2, False)
        # DCN contraction and gather (patch matrices rebuilt instead of stored)
        dcol = ops.dcn_im2col(x2, om, B, H, W, C)
        dwdcn = _z(wdcn)
        ops.gemm(ddcn, dcol, dwdcn, transA=True, transB=False, accumulate=True)
        ddcol = dcol                                        # reuse the buffer for the gradient
        ops.gemm(ddcn, wdcn, ddcol, transB=False)
        dxb, dom = ops.dcn_col2im(x2, om, ddcol, B, H, W, C)
        del ddcol, dcol
        dx = torch.empty_like(x2)
        ops.add2d(dxa, dxb, dx)
        dcol_i = None
        if implicit:
            from .convs import flip_weight_matrix
            wx32, wi32 = _pad_rows(wom_x), _pad_rows(wom_i)
            dwx32 = torch.zeros(32, 9 * C, device=dev)
            dwi32 = torch.zeros(32, 9 * Ci, device=dev)
            db32 = torch.zeros(32, device=dev)
            ops.conv3x3_wgrad(dom, x2, dwx32, B, H, W, accumulate=True, dbias=db32)
            ops.conv3x3_wgrad(dom, it2, dwi32, B, H, W, accumulate=True)
            dwom_x, dwom_i, dbom = dwx32[:27], dwi32[:27], db32[:27]
            dom3 = dom.view(B, H * W, 32)
            ops.conv3x3_gemm(dom3, flip_weight_matrix(wx32, C), dx, B, H, W, accumulate=True)
            if ctx.needs_input_grad[1]:
                ops.conv3x3_gemm(dom3, flip_weight_matrix(wi32, Ci), dit, B, H, W, accumulate=True)
        else:
            dbom = torch.empty(27, device=dev)
            ops.colsum(dom, dbom)
            colx = ops.im2col(x2, B, H, W, C, 3, 3, 1, 1)
            dwom_x, dwom_i = _z(wom_x), _z(wom_i)
            ops.gemm(dom, colx, dwom_x, transA=True, transB=False, accumulate=True)
            ops.gemm(dom, col_i, dwom_i, transA=True, transB=False, accumulate=True)
            ops.gemm(dom, wom_x, colx, transB=False)            # colx buffer becomes d(colx)
            dxc = ops.col2im(colx, B, H, W, C, 3, 3, 1, 1).view(T, C)
            if ctx.needs_input_grad[2]:
                dcol_i = torch.empty_like(col_i)
                ops.gemm(dom, wom_i, dcol_i, transB=False)
            ops.add2d(dx, dxc, dx)
        return (dx.view(B, H * W, C), dit.view(B, H * W, -1), dcol_i, dwom_x, dwom_i, dbom, dwdcn, dwg0, dwg2, dwb0,
                dwb2, None, None, None)


class SFT_layer(nn.Module):
    def __init__(self, channels_in, channels_out):
        super().__init__()
        self.conv_gamma = nn.Sequential(nn.Conv2d(channels_in, channels_out, 1, 1, 0, bias=False), nn.LeakyReLU(0.1, True),
                                        nn.Conv2d(channels_out, channels_out, 1, 1, 0, bias=False))
        self.conv_beta = nn.Sequential(nn.Conv2d(channels_in, channels_out, 1, 1, 0, bias=False), nn.LeakyReLU(0.1, True),
                                       nn.Conv2d(channels_out, channels_out, 1, 1, 0, bias=False))


class DGM(nn.Module):
    def __init__(self, channels_in, channels_out, kernel_size):
        super().__init__()
        self.channels_in, self.channels_out, self.kernel_size = channels_in, channels_out, kernel_size
        self.dcn = DCN_layer(channels_in, channels_out, kernel_size, padding=(kernel_size - 1) // 2, bias=False)
        self.sft = SFT_layer(channels_in, channels_out)
        self.relu = nn.LeakyReLU(0.1, True)

    def forward_tokens(self, x, it, col_i, H, W, slope=1.0):
        C = self.channels_in
        wom = self.dcn.conv_offset_mask.weight                                    # [27, 2C, 3, 3]
        wom_x = wom[:, :C].permute(0, 2, 3, 1).reshape(27, -1)
        wom_i = wom[:, C:].permute(0, 2, 3, 1).reshape(27, -1)
        s = self.sft

        def m(conv):
            return conv.weight.view(conv.weight.shape[0], -1)
        return DGMFn.apply(x, it, col_i, wom_x, wom_i, self.dcn.conv_offset_mask.bias, conv_weight_matrix(self.dcn.weight),
                           m(s.conv_gamma[0]), m(s.conv_gamma[2]), m(s.conv_beta[0]), m(s.conv_beta[2]), H, W, slope)


class DGB(nn.Module):
    def __init__(self, conv, n_feat, kernel_size):
        super().__init__()
        self.dgm1 = DGM(n_feat, n_feat, kernel_size)
        self.dgm2 = DGM(n_feat, n_feat, kernel_size)
        self.conv1 = conv(n_feat, n_feat, kernel_size)
        self.conv2 = conv(n_feat, n_feat, kernel_size)
        self.relu = nn.LeakyReLU(0.1, True)

    def forward_tokens(self, x, it, col_i, H, W):
        out = self.dgm1.forward_tokens(x, it, col_i, H, W, slope=0.1)             # relu(dgm1(x))  :79
        out = conv_tokens(out, self.conv1, H, W, ops.ACT_LRELU, 0.1)              # relu(conv1)    :80
        out = self.dgm2.forward_tokens(out, it, col_i, H, W, slope=0.1)           # relu(dgm2)     :81
        return conv_tokens(out, self.conv2, H, W, residual=x)                     # conv2 + x      :82


class DGG(nn.Module):
    def __init__(self, conv, n_feat, kernel_size, n_blocks):
        super().__init__()
        self.n_blocks = n_blocks
        body = [DGB(conv, n_feat, kernel_size) for _ in range(n_blocks)]
        body.append(conv(n_feat, n_feat, kernel_size))
        self.body = nn.Sequential(*body)

    def forward_tokens(self, x, it, col_i, H, W):
        res = x
        for i in range(self.n_blocks):
            res = self.body[i].forward_tokens(res, it, col_i, H, W)
        return conv_tokens(res, self.body[-1], H, W, residual=x)


def uformer_inter_to_map(inter, n_feats, H, W):
    """The adapter this package DEFINES for the Uformer-encoder + DGRN pairing (no counterpart in the reference, which
    crashes on it): the L per-band bottleneck features [B, 64, 448] (an 8 x 8 token grid, 448 channels) become the
    [B, n_feats, H, W] degradation map DGRN modulates with by (1) averaging the bands, (2) keeping the first n_feats
    channels, (3) nearest-neighbour upsampling of the 8 x 8 grid to H x W.  No parameters, so DGRN's state_dict stays the
    reference's; plain torch ops on 64 x 448 values per image (nothing here is on the hot path)."""
    m = torch.stack(list(inter), 0).mean(0)                                   # [B, 64, 448]
    B, N, C = m.shape
    g = int(round(N ** 0.5))
    assert g * g == N and C >= n_feats and H % g == 0 and W % g == 0
    m = m[:, :, :n_feats].transpose(1, 2).reshape(B, n_feats, g, g)
    return m.repeat_interleave(H // g, 2).repeat_interleave(W // g, 3).contiguous()


class DGRN(nn.Module):
    def __init__(self, opt, conv=default_conv):
        super().__init__()
        self.n_groups = 5
        n_blocks = 5
        if opt.encoder_type == 'ResNet':
            n_feats = opt.encoder_dim // 4
        elif opt.encoder_type == 'ViT':
            n_feats = opt.encoder_dim
        elif opt.encoder_type == 'Uformer':
            # The reference leaves n_feats undefined for this pairing (decoder_DGRN.py:120-129 -> UnboundLocalError)
            # and hands DGRN a tuple of L token tensors [B, 64, 448] instead of a [B, C, H, W] map: there is nothing to be
            # faithful to.  BASELINE configs[3] names the pairing, so it is DEFINED here, parameter-free and documented
            # (DESIGN.md section 0): n_feats = encoder_dim // 4 as for the ResNet encoder, and the degradation map is
            # uformer_inter_to_map(inter) below.  The state_dict is that of the ResNet-encoder DGRN.
            n_feats = opt.encoder_dim // 4
        if n_feats % 4:
            raise NotImplementedError(f'freqair: DGRN n_feats={n_feats} must be a multiple of 4 (128-bit NHWC gathers); '
                                      'use --encoder_dim 64 with the ViT encoder as the ViT runs of the reference authors do')
        self.n_feats = n_feats
        kernel_size = 3
        self.head = nn.Sequential(conv(3, n_feats, kernel_size))
        body = [DGG(default_conv, n_feats, kernel_size, n_blocks) for _ in range(self.n_groups)]
        body.append(conv(n_feats, n_feats, kernel_size))
        self.body = nn.Sequential(*body)
        self.tail = nn.Sequential(conv(n_feats, 3, kernel_size))

    def forward(self, x, inter):
        B, _, H, W = x.shape
        if isinstance(inter, (tuple, list)):               # Uformer encoder: L per-band token tensors [B, 64, 448]
            inter = uformer_inter_to_map(inter, self.n_feats, H, W)
        it = getattr(inter, '_fa_tokens', None)
        if it is None:
            it = NchwToTokensFn.apply(inter)
        # explicit path only: the patch matrix of the degradation map, shared by all 50 offset convolutions
        implicit = ops.conv3x3_eligible(H, W, self.n_feats, 32) and ops.conv3x3_eligible(H, W, it.shape[-1], 32)
        col_i = None if implicit else Im2colFn.apply(it, H, W)
        t = NchwToTokensFn.apply(x)
        h = conv_tokens(t, self.head[0], H, W)
        res = h
        for i in range(self.n_groups):
            res = self.body[i].forward_tokens(res, it, col_i, H, W)
        res = conv_tokens(res, self.body[-1], H, W, residual=h)
        out = conv_tokens(res, self.tail[0], H, W)
        return TokensToNchwFn.apply(out, H, W)
