"""Oracle: ResNet / ViT encoders, DGRN decoder with DCNv2 + SFT, MoCo wrapper and the
``train.py`` step body (test infrastructure; see oracle/__init__.py).
"""
import torch
import torch.nn.functional as F

from . import freq
from .uformer import lin, layer_norm


# ----------------------------------------------------------------------------- DCNv2
def modulated_deform_conv2d(x, offset, mask, weight, bias=None, pad=1):
    """DCNv2, stride 1, dilation 1, groups 1, deformable_groups 1, 3x3.

    **parity unpinned**: the reference's call is commented out and replaced by
    ``assert False`` (net/utils/deform_conv.py:64-67); mmcv is absent and has no
    pinned version.  This restates the published definition (Zhu et al. 2019, as
    implemented by mmcv's ``modulated_deformable_im2col`` and torchvision's
    ``deform_conv2d``): for tap k=(ki,kj), output pixel (y,x) samples the input
    bilinearly at (y - pad + ki + offset[2k], x - pad + kj + offset[2k+1]) with
    zeros outside the image, scales by mask[k], then contracts with ``weight``.
    """
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    ys = torch.arange(H, dtype=x.dtype).view(1, H, 1)
    xs = torch.arange(W, dtype=x.dtype).view(1, 1, W)
    cols = []
    for k in range(kh * kw):
        ki, kj = divmod(k, kw)
        py = ys - pad + ki + offset[:, 2 * k]
        px = xs - pad + kj + offset[:, 2 * k + 1]
        y0, x0 = torch.floor(py), torch.floor(px)
        ly, lx = py - y0, px - x0
        val = 0
        for dy, wy in ((0, 1 - ly), (1, ly)):
            for dx, wx in ((0, 1 - lx), (1, lx)):
                yy, xx = (y0 + dy).long(), (x0 + dx).long()
                ok = ((yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)).to(x.dtype)
                idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).view(B, 1, -1).expand(-1, C, -1)
                g = x.flatten(2).gather(2, idx).view(B, C, H, W)
                val = val + g * (wy * wx * ok).unsqueeze(1)
        cols.append(val * mask[:, k].unsqueeze(1))
    col = torch.stack(cols, 2)                                # [B, C, 9, H, W]
    out = torch.einsum('bckhw,ock->bohw', col, weight.flatten(2))
    return out if bias is None else out + bias.view(1, -1, 1, 1)


def dcn_layer(sd, p, x, inter):
    """DCN_layer.forward (deform_conv.py:56-67): offset/mask conv then DCNv2."""
    o = F.conv2d(torch.cat([x, inter], 1), sd[p + '.conv_offset_mask.weight'], sd[p + '.conv_offset_mask.bias'], padding=1)
    o1, o2, m = torch.chunk(o, 3, 1)
    return modulated_deform_conv2d(x, torch.cat((o1, o2), 1), torch.sigmoid(m), sd[p + '.weight'], sd.get(p + '.bias'))


def sft_layer(sd, p, x, inter):
    """SFT_layer.forward (decoder_DGRN.py:49-57)."""
    def mlp(q):
        return F.conv2d(F.leaky_relu(F.conv2d(inter, sd[f'{q}.0.weight']), 0.1), sd[f'{q}.2.weight'])
    return x * mlp(p + '.conv_gamma') + mlp(p + '.conv_beta')


def dgm(sd, p, x, inter):
    """DGM.forward (decoder_DGRN.py:22-32)."""
    return x + dcn_layer(sd, p + '.dcn', x, inter) + sft_layer(sd, p + '.sft', x, inter)


def conv3(sd, p, x):
    return F.conv2d(x, sd[p + '.weight'], sd[p + '.bias'], padding=1)


def dgb(sd, p, x, inter):
    """DGB.forward (decoder_DGRN.py:73-84)."""
    o = F.leaky_relu(dgm(sd, p + '.dgm1', x, inter), 0.1)
    o = F.leaky_relu(conv3(sd, p + '.conv1', o), 0.1)
    o = F.leaky_relu(dgm(sd, p + '.dgm2', o, inter), 0.1)
    return conv3(sd, p + '.conv2', o) + x


def dgrn_forward(sd, p, x, inter, n_groups=5, n_blocks=5):
    """DGRN.forward (decoder_DGRN.py:144-158)."""
    x = conv3(sd, p + 'head.0', x)
    res = x
    for g in range(n_groups):
        r = res
        for b in range(n_blocks):
            r = dgb(sd, f'{p}body.{g}.body.{b}', r, inter)
        res = conv3(sd, f'{p}body.{g}.body.{n_blocks}', r) + res                  # DGG.forward :99-110
    res = conv3(sd, f'{p}body.{n_groups}', res) + x
    return conv3(sd, p + 'tail.0', res)


# ----------------------------------------------------------------------------- ResNet encoder
def batch_norm(sd, p, x, training, bn_stats=None):
    if training:
        mean, var = x.mean((0, 2, 3)), x.var((0, 2, 3), unbiased=False)
        if bn_stats is not None:
            bn_stats[p] = (mean.detach(), x.var((0, 2, 3), unbiased=True).detach())
    else:
        mean, var = sd[p + '.running_mean'], sd[p + '.running_var']
    y = (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + 1e-5)
    return y * sd[p + '.weight'][None, :, None, None] + sd[p + '.bias'][None, :, None, None]


def res_block(sd, p, x, stride, training, bn_stats=None):
    """ResBlock (encoder_ResNet.py:4-20)."""
    y = F.conv2d(x, sd[p + '.backbone.0.weight'], stride=stride, padding=1)
    y = F.leaky_relu(batch_norm(sd, p + '.backbone.1', y, training, bn_stats), 0.1)
    y = batch_norm(sd, p + '.backbone.4', F.conv2d(y, sd[p + '.backbone.3.weight'], padding=1), training, bn_stats)
    s = batch_norm(sd, p + '.shortcut.1', F.conv2d(x, sd[p + '.shortcut.0.weight'], stride=stride), training, bn_stats)
    return F.leaky_relu(y + s, 0.1)


def resnet_encoder_forward(sd, p, x, training=False, bn_stats=None):
    """ResNetEncoder.forward (encoder_ResNet.py:42-47) -> (fea, [out], inter)."""
    inter = res_block(sd, p + 'E_pre', x, 1, training, bn_stats)
    f = res_block(sd, p + 'E.0', inter, 2, training, bn_stats)
    f = res_block(sd, p + 'E.1', f, 2, training, bn_stats).mean((2, 3))
    out = lin(sd, p + 'mlp.2', F.leaky_relu(lin(sd, p + 'mlp.0', f), 0.1))
    return f, [out], inter


# ----------------------------------------------------------------------------- ViT encoder
def vit_attention(sd, p, x, heads, decompose_type):
    """Attention.forward (encoder_ViT.py:76-98); dropout inactive (eval / p=0)."""
    B, N, _ = x.shape
    qkv = F.linear(x, sd[p + '.to_qkv.weight']).chunk(3, -1)
    q, k, v = (t.view(B, N, heads, -1).transpose(1, 2) for t in qkv)
    attn = (q @ k.transpose(-1, -2) * q.shape[-1] ** -0.5).softmax(-1)
    if decompose_type != 'none':
        lamb = sd[p + '.lamb']                               # [nb, 1 or B, heads]
        if decompose_type == 'DC':
            bands = freq.decompose(attn, 'frequency_decompose_dc', 0.5)
        else:
            nb = int(decompose_type.split('_')[0])
            bands = freq.decompose(attn, 'frequency_decompose', 1.0 / nb)
        attn = attn + (bands * lamb[:, :, :, None, None]).sum(0)
    out = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    return lin(sd, p + '.to_out.0', out)


def vit_encoder_forward(sd, p, x, encoder_dim, out_channels=3, decompose_type='none', depth=12, heads=12,
                        patch=16, training=False, bn_stats=None):
    """ViTEncoder.forward (encoder_ViT.py:182-203) -> (fea, [out], inter); dropout off."""
    B, C, H, W = x.shape
    t = x.view(B, C, H // patch, patch, W // patch, patch).permute(0, 2, 4, 3, 5, 1).reshape(B, -1, patch * patch * C)
    t = layer_norm(sd, p + 'to_patch_embedding.1', t)
    t = layer_norm(sd, p + 'to_patch_embedding.3', lin(sd, p + 'to_patch_embedding.2', t))
    t = t + sd[p + 'pos_embedding'][:, :t.shape[1]]
    for l in range(depth):
        q = f'{p}transformer.layers.{l}'
        t = vit_attention(sd, q + '.0.fn', layer_norm(sd, q + '.0.norm', t), heads, decompose_type) + t
        h = F.gelu(lin(sd, q + '.1.fn.net.0', layer_norm(sd, q + '.1.norm', t)))
        t = lin(sd, q + '.1.fn.net.3', h) + t
    t = lin(sd, p + 'mlp_head.1', layer_norm(sd, p + 'mlp_head.0', t))
    inter = t.reshape(-1, encoder_dim, H, W)
    inter = F.leaky_relu(batch_norm(sd, p + 'norm.0', inter, training, bn_stats), 0.1)
    fea = inter.mean((2, 3))
    out = lin(sd, p + 'mlp.2', F.leaky_relu(lin(sd, p + 'mlp.0', fea), 0.1))
    return fea, [out], inter


# ----------------------------------------------------------------------------- MoCo + train step
def moco_logits(q, k, queue, T=0.07):
    """MoCo.forward logits (moco.py:127-156) for per-band lists q, k and queue [L, dim, K]."""
    logits = []
    for i in range(len(q)):
        qi, ki = F.normalize(q[i], dim=1), F.normalize(k[i], dim=1)
        l_pos = (qi * ki).sum(1, keepdim=True)
        l_neg = qi @ queue[i]
        logits.append(torch.cat([l_pos, l_neg], 1) / T)
    return logits


def momentum_update(sd, q_prefix, k_prefix, names, m=0.999):
    """_momentum_update_key_encoder (moco.py:45-50) on parameter names (buffers excluded)."""
    for n in names:
        sd[k_prefix + n] = sd[k_prefix + n] * m + sd[q_prefix + n].detach() * (1.0 - m)


def adam_step(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (train.py:63,96; no weight decay, no amsgrad)."""
    m = m * b1 + g * (1 - b1)
    v = v * b2 + g * g * (1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / bc2 ** 0.5) + eps
    return p - (lr / bc1) * m / denom, m, v


def airnet_uformer_forward(sd, xq, xk, training, method='all_3_bands', L=3, msa='freq', encoder_dim=256,
                           dp=None, bn_stats=None, param_names=None):
    """AirNet.forward for Uformer encoder + Uformer decoder (net/model.py:59-71 over moco.py:115-170).

    training: returns (restored, logits, k_out) after the in-place momentum update of the
    ``E.E.encoder_k.*`` entries of ``sd``; eval: returns restored.
    """
    from . import uformer as U
    q_pre, k_pre = 'E.E.encoder_q.', 'E.E.encoder_k.'
    _, q, inter = U.encoder_forward(sd, q_pre, xq, L, msa, encoder_dim, training, dp, bn_stats)
    if not training:
        return U.decoder_forward(sd, 'R.R.', xq, inter, method, dp)
    with torch.no_grad():
        momentum_update(sd, q_pre, k_pre, param_names)
        _, k, _ = U.encoder_forward(sd, k_pre, xk, L, msa, encoder_dim, True, dp, bn_stats)
    logits = moco_logits(q, k, sd['E.E.queue'])
    restored = U.decoder_forward(sd, 'R.R.', xq, inter, method, dp)
    return restored, logits, k


def airnet_dgrn_forward(sd, xq, xk, training, encoder='ResNet', encoder_dim=256, decompose_type='none', bn_stats=None,
                        param_names=None):
    """AirNet.forward for the ResNet or ViT encoder + DGRN (net/model.py:59-71 over moco.py:115-170).

    The reference's MoCo loops ``range(opt.L)`` over the 1-element ``[out]`` these encoders return and dies with
    IndexError (moco.py:127-128 vs encoder_ResNet.py:47, encoder_ViT.py:203); the one deviation, shared by the product
    and by the golden generator (tools/make_golden_dgrn.py), is ``num_losses = len(out)``.  Dropout is off (p = 0).
    training: returns (restored, logits, k_out) after the in-place momentum update of the ``E.E.encoder_k.*`` entries.
    """
    q_pre, k_pre = 'E.E.encoder_q.', 'E.E.encoder_k.'

    def enc(p, x, train):
        if encoder == 'ResNet':
            return resnet_encoder_forward(sd, p, x, train, bn_stats)
        return vit_encoder_forward(sd, p, x, encoder_dim, decompose_type=decompose_type, training=train, bn_stats=bn_stats)
    _, q, inter = enc(q_pre, xq, training)
    if not training:
        return dgrn_forward(sd, 'R.R.', xq, inter)
    with torch.no_grad():
        momentum_update(sd, q_pre, k_pre, param_names)
        _, k, _ = enc(k_pre, xk, True)
    logits = moco_logits(q, k, sd['E.E.queue'])
    restored = dgrn_forward(sd, 'R.R.', xq, inter)
    return restored, logits, k
