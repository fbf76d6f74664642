"""Oracle: Uformer encoder / decoder (test infrastructure; see oracle/__init__.py).

Functional fp32 restatement of ``net/encoder_Uformer.py`` and
``net/decoder_Uformer.py`` for the configurations that run at reference HEAD
(SURVEY.md §8c): ``degradation_embedding_method`` = ``all_<n>_bands`` or
``all_DC``, ``frequency_decompose_type`` = ``none``, ``encoder_msa_type`` in
{``freq``, ``origin``}.  ``sd`` maps the reference's parameter names to
tensors; ``p`` is the key prefix of the module being evaluated.

``dp`` (optional) maps a block prefix to a per-sample DropPath scale vector
``[B']`` (0 or 1/keep) applied to both residual branches of that block
(the reference draws two independent masks per block; parity runs use
eval mode or pass explicit scales as ``(s_attn, s_mlp)``).
"""
import math

import torch
import torch.nn.functional as F

from . import freq

WIN = 8
DEC_DEPTHS = [2, 2, 8, 8, 2, 8, 8, 2, 2]          # decoder_Uformer.py:837
ENC_DEPTHS = [2, 2, 2, 2, 2]                      # encoder_Uformer.py:748 (first five are built)
HEADS = [1, 2, 4, 8, 16, 16, 8, 4, 2]


def lin(sd, p, x):
    return F.linear(x, sd[p + '.weight'], sd.get(p + '.bias'))


def layer_norm(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def rel_index(win=WIN):
    """relative_position_index buffer (decoder_Uformer.py:200-211)."""
    c = torch.stack(torch.meshgrid(torch.arange(win), torch.arange(win), indexing='ij')).flatten(1)
    r = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
    r[:, :, 0] += win - 1
    r[:, :, 1] += win - 1
    r[:, :, 0] *= 2 * win - 1
    return r.sum(-1)


def shift_mask(H, W, win=WIN, shift=WIN // 2):
    """SW-MSA additive mask [nW, 64, 64] of 0 / -100 (decoder_Uformer.py:634-651)."""
    m = torch.zeros(1, H, W, 1)
    cnt = 0
    for hs in (slice(0, -win), slice(-win, -shift), slice(-shift, None)):
        for ws in (slice(0, -win), slice(-win, -shift), slice(-shift, None)):
            m[:, hs, ws, :] = cnt
            cnt += 1
    mw = partition(m, win).view(-1, win * win)
    d = mw.unsqueeze(1) - mw.unsqueeze(2)
    return torch.where(d != 0, torch.tensor(-100.0), torch.tensor(0.0))


def partition(x, win=WIN):
    """[B,H,W,C] -> [B*nW, win, win, C] (decoder_Uformer.py:387-398)."""
    B, H, W, C = x.shape
    x = x.view(B, H // win, win, W // win, win, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, win, win, C)


def reverse(w, H, W, win=WIN):
    """inverse of partition (decoder_Uformer.py:400-409)."""
    B = w.shape[0] // ((H // win) * (W // win))
    x = w.view(B, H // win, W // win, win, win, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def leff(sd, p, x):
    """LeFF (leff.py:92-117): Linear+GELU, depthwise 3x3 + GELU, Linear."""
    B, L, _ = x.shape
    hh = int(math.sqrt(L))
    h = F.gelu(lin(sd, p + '.linear1.0', x))
    h = h.transpose(1, 2).reshape(B, -1, hh, hh)
    h = F.gelu(F.conv2d(h, sd[p + '.conv.0.weight'], sd[p + '.conv.0.bias'], padding=1, groups=h.shape[1]))
    h = h.flatten(2).transpose(1, 2)
    return lin(sd, p + '.linear2.0', h)


def qkv_heads(sd, p, x, heads):
    """LinearProjection (decoder_Uformer.py:98-125) -> q,k,v [B_, heads, N, hd]."""
    B_, N, C = x.shape
    q = lin(sd, p + '.to_q', x).reshape(B_, N, heads, C // heads).permute(0, 2, 1, 3)
    kv = lin(sd, p + '.to_kv', x).reshape(B_, N, 2, heads, C // heads).permute(2, 0, 3, 1, 4)
    return q, kv[0], kv[1]


def embed_lamb(sd, p, i, inter_i):
    """lambda predictor for band i (decoder_Uformer.py:178-193, 280-284) -> [B,1,heads]."""
    e = lin(sd, f'{p}.mlp_head.{i}.1', layer_norm(sd, f'{p}.mlp_head.{i}.0', inter_i))   # [B,64,heads]
    e = e.mean(1, keepdim=True)                                                      # AdaptiveAvgPool1d over tokens
    e = F.leaky_relu(lin(sd, f'{p}.mlp.{i}.0', e), 0.1)
    return lin(sd, f'{p}.mlp.{i}.2', e)


def dec_window_attention(sd, p, x, heads, mask, all_inter, method, num_win):
    """WindowAttention.forward of the decoder (decoder_Uformer.py:235-299)."""
    B_, N, C = x.shape
    q, k, v = qkv_heads(sd, p + '.qkv', x, heads)
    attn = (q * (C // heads) ** -0.5) @ k.transpose(-2, -1)
    bias = sd[p + '.relative_position_bias_table'][rel_index().view(-1)].view(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    attn = attn.softmax(-1)
    if method is not None:
        if method.endswith('bands'):
            nb = int(method.split('_')[-2])
            bands = freq.decompose(attn, 'frequency_decompose_1', 1.0 / (nb - 1))
        else:                                               # all_DC
            nb = 2
            bands = freq.decompose(attn, 'frequency_decompose_dc', 0.5)
        for i in range(1, nb):
            lam = embed_lamb(sd, p, i, all_inter[i])        # [B,1,heads]
            band = bands[i].view(-1, num_win, heads, N, N) * lam[:, :, :, None, None]
            attn = attn + band.view(-1, heads, N, N)
    out = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return lin(sd, p + '.proj', out)


def _dp(dp, p, which, x):
    if dp is None or p not in dp:
        return x
    s = dp[p][which]
    return x * s.view(-1, 1, 1)


def dec_block(sd, p, x, heads, shift, all_inter, method, dp=None):
    """LeWinTransformerBlock.forward of the decoder (decoder_Uformer.py:618-756), plain path."""
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    mask = shift_mask(H, W) if shift > 0 else None
    y = layer_norm(sd, p + '.norm1', x).view(B, H, W, C)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    win = partition(y).view(-1, WIN * WIN, C)
    a = dec_window_attention(sd, p + '.attn', win, heads, mask, all_inter, method, (H // WIN) * (W // WIN))
    y = reverse(a.view(-1, WIN, WIN, C), H, W)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    x = x + _dp(dp, p, 0, y.view(B, L, C))
    return x + _dp(dp, p, 1, leff(sd, p + '.mlp', layer_norm(sd, p + '.norm2', x)))


def layer(sd, p, x, depth, heads, blockfn, **kw):
    """BasicUformerLayer (decoder_Uformer.py:761-832): even blocks W-MSA, odd blocks shifted."""
    H = int(math.sqrt(x.shape[1]))
    for i in range(depth):
        shift = WIN // 2 if (i % 2 == 1 and H > WIN) else 0
        x = blockfn(sd, f'{p}.blocks.{i}', x, heads, shift, **kw)
    return x


def conv_tokens(sd, p, x, **kw):
    """tokens -> NCHW -> Conv2d -> tokens (Downsample, decoder_Uformer.py:423-430)."""
    B, L, C = x.shape
    H = int(math.sqrt(L))
    y = F.conv2d(x.transpose(1, 2).reshape(B, C, H, H), sd[p + '.weight'], sd[p + '.bias'], **kw)
    return y.flatten(2).transpose(1, 2)


def deconv_tokens(sd, p, x):
    """Upsample: ConvTranspose2d k2 s2 on tokens (decoder_Uformer.py:443-449)."""
    B, L, C = x.shape
    H = int(math.sqrt(L))
    y = F.conv_transpose2d(x.transpose(1, 2).reshape(B, C, H, H), sd[p + '.weight'], sd[p + '.bias'], stride=2)
    return y.flatten(2).transpose(1, 2)


def input_proj(sd, p, x):
    """InputProj: conv3x3 + LeakyReLU(0.01) -> tokens (decoder_Uformer.py:453-472)."""
    y = F.leaky_relu(F.conv2d(x, sd[p + '.proj.0.weight'], sd[p + '.proj.0.bias'], padding=1), 0.01)
    return y.flatten(2).transpose(1, 2)


def decoder_forward(sd, p, x, all_inter, method='all_3_bands', dp=None):
    """UformerDecoder.forward (decoder_Uformer.py:1117-1171) -> restored [B,3,H,W].

    ``all_inter``: the encoder's per-band token features, tuple of L x [B,64,448].
    """
    kw = dict(all_inter=all_inter, method=method, dp=dp)
    y = input_proj(sd, p + 'input_proj', x)
    skips = []
    for i in range(4):
        y = layer(sd, f'{p}encoderlayer_{i}', y, DEC_DEPTHS[i], HEADS[i], dec_block, **kw)
        skips.append(y)
        y = conv_tokens(sd, f'{p}dowsample_{i}.conv.0', y, stride=2, padding=1)
    y = layer(sd, p + 'bottleneck_0', y, DEC_DEPTHS[4], HEADS[4], dec_block, **kw)
    y = layer(sd, p + 'bottleneck_1', y, DEC_DEPTHS[4], HEADS[4], dec_block, **kw)
    for n, i in enumerate(reversed(range(4))):
        y = deconv_tokens(sd, f'{p}upsample_{i}.deconv.0', y)
        y = torch.cat([y, skips[i]], -1)
        y = layer(sd, f'{p}decoderlayer_{i}', y, DEC_DEPTHS[5 + n], HEADS[5 + n], dec_block, **kw)
    B, L, C = y.shape
    H = int(math.sqrt(L))
    out = F.conv2d(y.transpose(1, 2).reshape(B, C, H, H), sd[p + 'output_proj.proj.0.weight'],
                   sd[p + 'output_proj.proj.0.bias'], padding=1)
    return x + out


# ----------------------------------------------------------------------------- encoder
def joint_attention_core(q, k, v, tables, mask, L, kind):
    """Attention core of FrequencyWindowAttention.forward (encoder_Uformer.py:259-300) on per-head
    q, k, v [(l bnw), heads, 64, hd]; ``tables``: the L*L bias tables [225, heads] -> [(l bnw), heads, 64, hd]."""
    B_, heads, N, hd = q.shape

    def join(t):                                             # (l bnw) h t d -> bnw h (l t) d
        return t.reshape(L, B_ // L, heads, N, hd).permute(1, 2, 0, 3, 4).reshape(B_ // L, heads, L * N, hd)
    q, k, v = join(q), join(k), join(v)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    ri = rel_index().view(-1)
    bias = torch.stack([tables[i][ri].view(N, N, -1).permute(2, 0, 1) for i in range(L * L)], 0)   # [(l1 l2), h, t1, t2]
    bias = bias.view(L, L, heads, N, N).permute(2, 0, 3, 1, 4).reshape(1, heads, L * N, L * N)
    same = torch.eye(L).repeat_interleave(N, 0).repeat_interleave(N, 1)
    mfreq = (1 - same) * -100.0 if kind == 'intra' else same * -100.0          # :246-254
    attn = attn + bias + mfreq
    if mask is not None:
        nW = mask.shape[0]
        big = mask.repeat(1, L, L)                            # nW (l1 t1) (l2 t2)
        attn = (attn.view(-1, nW, heads, L * N, L * N) + big.unsqueeze(1).unsqueeze(0)).view(-1, heads, L * N, L * N)
    attn = attn.softmax(-1)
    o = attn @ v                                             # bnw h (l t) d
    return o.view(B_ // L, heads, L, N, hd).permute(2, 0, 1, 3, 4).reshape(B_, heads, N, hd)


def enc_freq_attention(sd, p, x, heads, mask, L, kind):
    """FrequencyWindowAttention.forward (encoder_Uformer.py:256-310): joint attention over
    the L band copies of one window; ``kind`` 'intra' masks off-band pairs, 'inter' same-band pairs."""
    B_, N, C = x.shape                                       # B_ = (l b nw)
    q, k, v = qkv_heads(sd, p + '.qkv', x, heads)
    tables = [sd[f'{p}.relative_position_bias_table.{i}'] for i in range(L * L)]
    o = joint_attention_core(q, k, v, tables, mask, L, kind)
    return lin(sd, p + '.proj', o.transpose(1, 2).reshape(B_, N, C))


def enc_origin_attention(sd, p, x, heads, mask):
    """WindowAttention.forward of the encoder (encoder_Uformer.py:152-183)."""
    B_, N, C = x.shape
    q, k, v = qkv_heads(sd, p + '.qkv', x, heads)
    attn = (q * (C // heads) ** -0.5) @ k.transpose(-2, -1)
    attn = attn + sd[p + '.relative_position_bias_table'][rel_index().view(-1)].view(N, N, -1).permute(2, 0, 1)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    out = (attn.softmax(-1) @ v).transpose(1, 2).reshape(B_, N, C)
    return lin(sd, p + '.proj', out)


def enc_block(sd, p, x, heads, shift, L, msa, dp=None):
    """LeWinTransformerBlock.forward of the encoder (encoder_Uformer.py:597-682)."""
    B, T, C = x.shape
    H = W = int(math.sqrt(T))
    mask = shift_mask(H, W) if shift > 0 else None
    y = layer_norm(sd, p + '.norm1', x).view(B, H, W, C)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    win = partition(y).view(-1, WIN * WIN, C)
    if msa == 'origin':
        a = enc_origin_attention(sd, p + '.attn', win, heads, mask)
    else:
        a = enc_freq_attention(sd, p + '.attn_intra', win, heads, mask, L, 'intra')
        a = enc_freq_attention(sd, p + '.attn_inter', a, heads, mask, L, 'inter')
    y = reverse(a.view(-1, WIN, WIN, C), H, W)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    x = x + _dp(dp, p, 0, y.view(B, T, C))
    return x + _dp(dp, p, 1, leff(sd, p + '.mlp', layer_norm(sd, p + '.norm2', x)))


def encoder_trunk(sd, p, x, L=3, msa='freq', dp=None):
    """Uformer.forward of the encoder (encoder_Uformer.py:905-923): (L*B) images -> [(L*B), 64, 448]."""
    kw = dict(L=L, msa=msa, dp=dp)
    y = input_proj(sd, p + 'input_proj', x)
    for i in range(4):
        y = layer(sd, f'{p}encoderlayer_{i}', y, ENC_DEPTHS[i], HEADS[i], enc_block, **kw)
        y = conv_tokens(sd, f'{p}dowsample_{i}.conv.0', y, stride=2, padding=1)
    return layer(sd, p + 'conv', y, ENC_DEPTHS[4], HEADS[4], enc_block, **kw)


def encoder_head(sd, p, i, xi, encoder_dim, img, training, bn_stats=None):
    """Per-band contrastive head (encoder_Uformer.py:975-984): LN -> Linear(448 -> ed*256)
    -> raw reshape [B, ed, img, img] -> BatchNorm2d -> LeakyReLU(0.1) -> avgpool -> MLP."""
    f = lin(sd, f'{p}mlp_head.{i}.1', layer_norm(sd, f'{p}mlp_head.{i}.0', xi))
    f = f.reshape(f.shape[0], encoder_dim, img, img)
    bn = f'{p}norm.{i}.0'
    if training:
        mean = f.mean((0, 2, 3))
        var = f.var((0, 2, 3), unbiased=False)
        if bn_stats is not None:
            bn_stats[bn] = (mean.detach(), f.var((0, 2, 3), unbiased=True).detach())
    else:
        mean, var = sd[bn + '.running_mean'], sd[bn + '.running_var']
    f = (f - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + 1e-5)
    f = f * sd[bn + '.weight'][None, :, None, None] + sd[bn + '.bias'][None, :, None, None]
    f = F.leaky_relu(f, 0.1).mean((2, 3))
    f = F.leaky_relu(lin(sd, f'{p}mlp.{i}.0', f), 0.1)
    return lin(sd, f'{p}mlp.{i}.2', f)


def encoder_forward(sd, p, x, L=3, msa='freq', encoder_dim=256, training=False, dp=None, bn_stats=None):
    """UformerEncoder.forward (encoder_Uformer.py:959-986) -> (None, [out_i], inter tuple)."""
    B, _, img, _ = x.shape
    if L != 1:
        x = freq.decompose(x, 'frequency_decompose_1', 1.0 / (L - 1)).flatten(0, 1)     # (l b) c h w
    t = encoder_trunk(sd, p + 'uformer.', x, L, msa, dp)
    inter = tuple(t.view(L, B, *t.shape[1:]).unbind(0))
    out = [encoder_head(sd, p, i, inter[i], encoder_dim, img, training, bn_stats) for i in range(L)]
    return None, out, inter
