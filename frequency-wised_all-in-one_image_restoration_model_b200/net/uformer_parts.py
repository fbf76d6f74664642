"""Pieces shared by the Uformer encoder and decoder: parameter containers with the reference's
attribute names (so ``state_dict`` keys match net/encoder_Uformer.py and net/decoder_Uformer.py
one-to-one) whose ``forward`` enqueues libfreqair kernels through the block-level autograd nodes."""
import math

import torch
import torch.nn as nn

from .. import ops
from .convs import (ConvTokFn, InputProjFn, OutputProjFn, UpsampleCatFn, conv_weight_matrix,
                    deconv_weight_matrix)
from .utils.leff import LeFF

WIN = 8


def trunc_normal_(t, std=.02):
    return nn.init.trunc_normal_(t, std=std)


def relative_position_index(win=WIN):
    """int64 [64,64] buffer kept for checkpoint compatibility (decoder_Uformer.py:200-211); the kernels
    recompute the index arithmetically."""
    coords = torch.stack(torch.meshgrid(torch.arange(win), torch.arange(win), indexing='ij')).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += win - 1
    rel[:, :, 1] += win - 1
    rel[:, :, 0] *= 2 * win - 1
    return rel.sum(-1)


class LinearProjection(nn.Module):
    """to_q / to_kv parameter holder (decoder_Uformer.py:80-96)."""

    def __init__(self, dim, heads=8, dim_head=64, bias=True):
        super().__init__()
        inner = dim_head * heads
        self.heads = heads
        self.to_q = nn.Linear(dim, inner, bias=bias)
        self.to_kv = nn.Linear(dim, inner * 2, bias=bias)
        self.dim, self.inner_dim = dim, inner


class Downsample(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size=4, stride=2, padding=1):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_channel, out_channel, kernel_size=kernel_size, stride=stride,
                                            padding=padding))
        self.in_channel, self.out_channel = in_channel, out_channel
        self.k, self.s, self.p = kernel_size, stride, padding

    def forward(self, x):
        B, L, C = x.shape
        H = W = int(math.sqrt(L))
        c = self.conv[0]
        return ConvTokFn.apply(x, conv_weight_matrix(c.weight), c.bias, H, W, self.k, self.s, self.p,
                                  ops.ACT_NONE, 0.0, None, ops.BWD_BACKEND)


class Upsample(nn.Module):
    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.deconv = nn.Sequential(nn.ConvTranspose2d(in_channel, out_channel, kernel_size=2, stride=2))
        self.in_channel, self.out_channel = in_channel, out_channel

    def forward(self, x, skip=None):
        """Returns cat([deconv(x), skip], -1) when ``skip`` is given (the decoder always concatenates,
        decoder_Uformer.py:1157-1162), else deconv(x)."""
        B, L, C = x.shape
        H = W = int(math.sqrt(L))
        d = self.deconv[0]
        if skip is None:
            skip = x.new_zeros(B, 4 * L, 0)
        return UpsampleCatFn.apply(x, deconv_weight_matrix(d.weight), d.bias.repeat(4), skip, H, W)


class InputProj(nn.Module):
    def __init__(self, in_channel=3, out_channel=64, kernel_size=3, stride=1, norm_layer=None, act_layer=nn.LeakyReLU):
        super().__init__()
        self.proj = nn.Sequential(nn.Conv2d(in_channel, out_channel, kernel_size=3, stride=stride,
                                            padding=kernel_size // 2), act_layer(inplace=True))
        self.norm = norm_layer(out_channel) if norm_layer is not None else None
        self.in_channel, self.out_channel = in_channel, out_channel

    def forward(self, x):
        c = self.proj[0]
        y = InputProjFn.apply(x, conv_weight_matrix(c.weight), c.bias, self.proj[1].negative_slope)
        assert self.norm is None
        return y


class OutputProj(nn.Module):
    def __init__(self, in_channel=64, out_channel=3, kernel_size=3, stride=1, norm_layer=None, act_layer=None):
        super().__init__()
        self.proj = nn.Sequential(nn.Conv2d(in_channel, out_channel, kernel_size=3, stride=stride,
                                            padding=kernel_size // 2))
        assert act_layer is None and norm_layer is None
        self.norm = None
        self.in_channel, self.out_channel = in_channel, out_channel

    def forward(self, x, residual=None):
        """tokens -> [B, out, H, W] (+ residual image, fused)."""
        B, L, C = x.shape
        H = W = int(math.sqrt(L))
        c = self.proj[0]
        return OutputProjFn.apply(x, conv_weight_matrix(c.weight), c.bias, residual, H, W)


def draw_drop_path(x, drop_prob, training):
    """Per-sample DropPath scale vector (timm DropPath: Bernoulli(keep)/keep), or None."""
    if drop_prob == 0. or not training:
        return None
    keep = 1.0 - drop_prob
    r = torch.empty(x.shape[0], device=x.device, dtype=torch.float32).bernoulli_(keep)
    return r.div_(keep) if keep > 0 else r


def leff_params(mlp: LeFF):
    return (mlp.linear1[0].weight, mlp.linear1[0].bias, mlp.conv[0].weight, mlp.conv[0].bias,
            mlp.linear2[0].weight, mlp.linear2[0].bias)


def init_uformer_weights(m):
    """Uformer._init_weights (encoder_Uformer.py:885-892)."""
    if isinstance(m, nn.Linear):
        trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        if m.elementwise_affine:
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
