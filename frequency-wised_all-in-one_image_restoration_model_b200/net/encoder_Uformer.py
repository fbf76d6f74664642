"""UformerEncoder - drop-in for the reference's net/encoder_Uformer.py (``UformerEncoder(opt)``,
``forward(x) -> (None, [out_0..out_{L-1}], inter)``, identical state_dict keys) on libfreqair kernels.

Pipeline (encoder_Uformer.py:959-986): image -> L radial frequency bands (K1 band split at 128x128) ->
5-level LeWin encoder on the (l b) batch with intra- then inter-band joint window attention (K2 joint form)
-> per-band contrastive heads LN -> Linear(448 -> encoder_dim*256) -> raw reshape [B, encoder_dim, 128, 128]
-> BatchNorm2d -> LeakyReLU(0.1) -> global average pool -> MLP.  The head never materialises a second copy of
the [B, 256, 128, 128] activation: BN statistics, the activation and the pool are two streaming passes over the
GEMM output.
"""
import torch
import torch.nn as nn

from .. import ops
from .convs import BNHeadFn
from .lewin import EncoderBlockFn, layer_norm, linear
from .uformer_parts import (WIN, Downsample, InputProj, LinearProjection, draw_drop_path, init_uformer_weights,
                            leff_params, relative_position_index, trunc_normal_)
from .utils.frequency_decompose import FrequencyDecompose
from .utils.leff import LeFF


class WindowAttention(nn.Module):
    """'origin' MSA parameter holder (encoder_Uformer.py:103-150)."""

    def __init__(self, dim, win_size, num_heads):
        super().__init__()
        self.dim, self.win_size, self.num_heads = dim, win_size, num_heads
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * win_size[0] - 1) * (2 * win_size[1] - 1), num_heads))
        self.register_buffer('relative_position_index', relative_position_index(win_size[0]))
        trunc_normal_(self.relative_position_bias_table, std=.02)
        self.qkv = LinearProjection(dim, num_heads, dim // num_heads, bias=True)
        self.proj = nn.Linear(dim, dim)

    def tables(self):
        return self.relative_position_bias_table


class FrequencyWindowAttention(nn.Module):
    """intra / inter band attention parameter holder (encoder_Uformer.py:190-254): L*L bias tables."""

    def __init__(self, dim, win_size, num_heads, type=None, L=3):
        super().__init__()
        assert type in ('intra', 'inter'), 'Attention type error.'
        self.dim, self.win_size, self.num_heads, self.L, self.type = dim, win_size, num_heads, L, type
        self.relative_position_bias_table = nn.ParameterList([
            nn.Parameter(torch.zeros((2 * win_size[0] - 1) * (2 * win_size[1] - 1), num_heads)) for _ in range(L * L)])
        self.register_buffer('relative_position_index', relative_position_index(win_size[0]))
        for i in range(L * L):
            trunc_normal_(self.relative_position_bias_table[i], std=.02)
        self.qkv = LinearProjection(dim, num_heads, dim // num_heads, bias=True)
        self.proj = nn.Linear(dim, dim)
        n = win_size[0] * win_size[1]
        same = torch.eye(L).repeat_interleave(n, 0).repeat_interleave(n, 1)
        mask_freq = (1 - same) * -100.0 if type == 'intra' else same * -100.0
        # checkpoint-compatibility buffer only: the kernel regenerates the 0/-100 pattern from `kind`
        self.register_buffer('mask_freq', mask_freq.unsqueeze(0).unsqueeze(0))

    def tables(self):
        return torch.stack(list(self.relative_position_bias_table), 0)       # [L*L, 225, heads]


def _attn_params(a):
    return (a.tables(), a.qkv.to_q.weight, a.qkv.to_q.bias, a.qkv.to_kv.weight, a.qkv.to_kv.bias, a.proj.weight,
            a.proj.bias)


class LeWinTransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, win_size=8, shift_size=0, mlp_ratio=4., drop_path=0.,
                 encoder_msa_type=None, L=3):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        self.win_size, self.shift_size, self.mlp_ratio, self.L = win_size, shift_size, mlp_ratio, L
        if min(input_resolution) <= win_size:
            self.shift_size = 0
            self.win_size = min(input_resolution)
        assert self.win_size == WIN, 'freqair: 8x8 windows only (all reference geometries)'
        self.norm1 = nn.LayerNorm(dim)
        self.encoder_msa_type = encoder_msa_type
        ws = (self.win_size, self.win_size)
        if encoder_msa_type == 'origin':
            self.attn = WindowAttention(dim, ws, num_heads)
        elif encoder_msa_type == 'freq':
            self.attn_intra = FrequencyWindowAttention(dim, ws, num_heads, type='intra', L=L)
            self.attn_inter = FrequencyWindowAttention(dim, ws, num_heads, type='inter', L=L)
        else:
            assert False, 'MSA type error.'
        self.drop_path_prob = float(drop_path)
        self.drop_path = nn.Identity()
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = LeFF(dim, int(dim * mlp_ratio))
        self.forced_dp = None

    def forward(self, x, mask=None):
        assert mask is None
        LB, T, C = x.shape
        H = W = int(T ** 0.5)
        if self.forced_dp is not None:
            dp_a, dp_m = self.forced_dp
        else:
            dp_a = draw_drop_path(x, self.drop_path_prob, self.training)
            dp_m = draw_drop_path(x, self.drop_path_prob, self.training)
        if self.encoder_msa_type == 'origin':
            pa, pb = _attn_params(self.attn), (None,) * 7
            L = 1
        else:
            pa, pb = _attn_params(self.attn_intra), _attn_params(self.attn_inter)
            L = self.L
        cfg = (L, LB // L, H, W, self.num_heads, self.shift_size, self.encoder_msa_type)
        y = EncoderBlockFn.apply(cfg, x, dp_a, dp_m, self.norm1.weight, self.norm1.bias, *pa, *pb, self.norm2.weight,
                                 self.norm2.bias, *leff_params(self.mlp))
        return y, None, None


class BasicUformerLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, win_size, mlp_ratio, drop_path, encoder_msa_type, L):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.blocks = nn.ModuleList([
            LeWinTransformerBlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads, win_size=win_size,
                                  shift_size=0 if (i % 2 == 0) else win_size // 2, mlp_ratio=mlp_ratio,
                                  drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                  encoder_msa_type=encoder_msa_type, L=L)
            for i in range(depth)])

    def forward(self, x, mask=None):
        K = V = None
        for blk in self.blocks:
            x, K, V = blk(x, mask)
        return x, K, V


class Uformer(nn.Module):
    """Encoder half only (encoder_Uformer.py:746-923): input_proj, 4 x (layer + downsample), bottleneck `conv`."""

    def __init__(self, opt, img_size=128, in_chans=3, out_chans=3, depths=[2, 2, 2, 2, 2, 2, 2, 2, 2],
                 num_heads=[1, 2, 4, 8, 16, 16, 8, 4, 2], win_size=8, mlp_ratio=4., drop_path_rate=0.1, **kwargs):
        super().__init__()
        self.opt = opt
        embed_dim = opt.encoder_embed_dim
        L, msa = opt.L, opt.encoder_msa_type
        if 'attention_kv' in opt.degradation_embedding_method:
            raise NotImplementedError('freqair: attention_kv does not run at reference HEAD (decoder_Uformer.py:1124)')
        self.num_enc_layers = len(depths) // 2
        self.embed_dim, self.mlp_ratio, self.win_size, self.reso, self.in_chans = embed_dim, mlp_ratio, win_size, img_size, in_chans
        self.pos_drop = nn.Dropout(p=0.)
        enc_dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths[:self.num_enc_layers]))]
        conv_dpr = [drop_path_rate] * depths[4]

        def layer(mult, div, depth, heads, dpr):
            return BasicUformerLayer(dim=embed_dim * mult, input_resolution=(img_size // div, img_size // div),
                                     depth=depth, num_heads=heads, win_size=win_size, mlp_ratio=mlp_ratio,
                                     drop_path=dpr, encoder_msa_type=msa, L=L)
        self.input_proj = InputProj(in_channel=in_chans, out_channel=embed_dim, kernel_size=3, stride=1, act_layer=nn.LeakyReLU)
        self.encoderlayer_0 = layer(1, 1, depths[0], num_heads[0], enc_dpr[sum(depths[:0]):sum(depths[:1])])
        self.dowsample_0 = Downsample(embed_dim, embed_dim * 2)
        self.encoderlayer_1 = layer(2, 2, depths[1], num_heads[1], enc_dpr[sum(depths[:1]):sum(depths[:2])])
        self.dowsample_1 = Downsample(embed_dim * 2, embed_dim * 4)
        self.encoderlayer_2 = layer(4, 4, depths[2], num_heads[2], enc_dpr[sum(depths[:2]):sum(depths[:3])])
        self.dowsample_2 = Downsample(embed_dim * 4, embed_dim * 8)
        self.encoderlayer_3 = layer(8, 8, depths[3], num_heads[3], enc_dpr[sum(depths[:3]):sum(depths[:4])])
        self.dowsample_3 = Downsample(embed_dim * 8, embed_dim * 16)
        self.conv = layer(16, 16, depths[4], num_heads[4], conv_dpr)
        self.apply(init_uformer_weights)

    def forward(self, x, mask=None):
        y = self.input_proj(x)
        for lay, down in ((self.encoderlayer_0, self.dowsample_0), (self.encoderlayer_1, self.dowsample_1),
                          (self.encoderlayer_2, self.dowsample_2), (self.encoderlayer_3, self.dowsample_3)):
            y, _, _ = lay(y, mask=mask)
            y = down(y)
        y, _, _ = self.conv(y, mask=mask)
        return y


class UformerEncoder(nn.Module):
    def __init__(self, opt, img_size=128, in_chans=3, out_chans=3):
        super().__init__()
        self.opt = opt
        embed_dim = opt.encoder_embed_dim
        self.img_size = img_size
        if not opt.L == 1:
            self.preprocess_decompose = FrequencyDecompose('frequency_decompose_1', 1. / (opt.L - 1), img_size, img_size)
        self.uformer = Uformer(opt, img_size=img_size, in_chans=in_chans, out_chans=None, embed_dim=embed_dim)
        self.mlp_head = nn.ModuleList([nn.Sequential(nn.LayerNorm(embed_dim * 16),
                                                     nn.Linear(embed_dim * 16, opt.encoder_dim * 16 * 16))
                                       for _ in range(opt.L)])
        self.norm = nn.ModuleList([nn.Sequential(nn.BatchNorm2d(opt.encoder_dim), nn.LeakyReLU(0.1, True))
                                   for _ in range(opt.L)])
        self.avg = nn.ModuleList([nn.AdaptiveAvgPool2d(1) for _ in range(opt.L)])
        self.mlp = nn.ModuleList([nn.Sequential(nn.Linear(opt.encoder_dim, opt.encoder_dim), nn.LeakyReLU(0.1, True),
                                                nn.Linear(opt.encoder_dim, opt.encoder_dim)) for _ in range(opt.L)])

    def trunk(self, x):
        """x [B,3,H,W] -> tuple of L per-band token features [B, 64, 448] (the decoder's ``inter``)."""
        B = x.shape[0]
        L = self.opt.L
        if not L == 1:
            x = self.preprocess_decompose(x).flatten(0, 1)                  # l b c h w -> (l b) c h w
        t = self.uformer(x)
        return tuple(t.view(L, B, *t.shape[1:]).unbind(0))

    def head(self, i, xi):
        """Contrastive head of band i (encoder_Uformer.py:975-984)."""
        ed, S = self.opt.encoder_dim, self.img_size * self.img_size
        ln, fc = self.mlp_head[i][0], self.mlp_head[i][1]
        BB = ops.BWD_BACKEND
        f = linear(layer_norm(xi, ln.weight, ln.bias), fc.weight, fc.bias, bwd_backend=BB)            # [B, 64, ed*256]
        bn = self.norm[i][0]
        if self.training:
            bn.num_batches_tracked += 1
        pooled, _ = BNHeadFn.apply(f.reshape(f.shape[0], ed, S), bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                   self.training, 0.1, False)
        m = self.mlp[i]
        return linear(linear(pooled, m[0].weight, m[0].bias, ops.ACT_LRELU, 0.1, bwd_backend=BB), m[2].weight, m[2].bias,
                      bwd_backend=BB)

    def forward(self, x, mask=None):
        assert mask is None
        inter = self.trunk(x)
        out = [self.head(i, inter[i]) for i in range(self.opt.L)]
        return None, out, inter
