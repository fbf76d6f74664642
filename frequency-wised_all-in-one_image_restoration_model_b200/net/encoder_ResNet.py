"""ResNetEncoder - drop-in for the reference's net/encoder_ResNet.py (``ResNetEncoder(opt)``,
``forward(x) -> (fea, [out], inter)``, identical state_dict keys).  Internally everything stays in token
(NHWC) layout: each conv is a patch gather + GEMM, BatchNorm / LeakyReLU / the residual add are one fused
streaming kernel per ResBlock branch; ``inter`` is returned in the reference's NCHW layout and additionally
carries its token form (``inter._fa_tokens``) so the DGRN decoder can skip the round trip."""
import torch
from torch import nn

from .. import ops
from .convs import NchwToTokensFn, TokenMeanFn, TokensToNchwFn, bn_tokens, conv_tokens
from .lewin import linear


class ResBlock(nn.Module):
    def __init__(self, in_feat, out_feat, stride=1):
        super().__init__()
        self.backbone = nn.Sequential(
            nn.Conv2d(in_feat, out_feat, kernel_size=3, stride=stride, padding=1, bias=False),
            nn.BatchNorm2d(out_feat),
            nn.LeakyReLU(0.1, True),
            nn.Conv2d(out_feat, out_feat, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(out_feat),
        )
        self.shortcut = nn.Sequential(
            nn.Conv2d(in_feat, out_feat, kernel_size=1, stride=stride, bias=False),
            nn.BatchNorm2d(out_feat),
        )
        self.stride = stride

    def forward_tokens(self, x, H, W):
        """x [B, H*W, Cin] -> ([B, Ho*Wo, Cout], Ho, Wo)  (encoder_ResNet.py:19-20)."""
        tr = self.training
        Ho, Wo = (H - 1) // self.stride + 1, (W - 1) // self.stride + 1
        y = conv_tokens(x, self.backbone[0], H, W)
        y = bn_tokens(y, self.backbone[1], tr, slope=0.1)
        y = conv_tokens(y, self.backbone[3], Ho, Wo)
        s = conv_tokens(x, self.shortcut[0], H, W)
        s = bn_tokens(s, self.shortcut[1], tr)
        return bn_tokens(y, self.backbone[4], tr, slope=0.1, res=s), Ho, Wo


class ResNetEncoder(nn.Module):
    embedding_is_none = False

    def __init__(self, opt):
        super().__init__()
        self.dim = opt.encoder_dim
        self.E_pre = ResBlock(in_feat=3, out_feat=self.dim // 4, stride=1)
        self.E = nn.Sequential(
            ResBlock(in_feat=self.dim // 4, out_feat=self.dim // 2, stride=2),
            ResBlock(in_feat=self.dim // 2, out_feat=self.dim, stride=2),
            nn.AdaptiveAvgPool2d(1),
        )
        self.mlp = nn.Sequential(nn.Linear(self.dim, self.dim), nn.LeakyReLU(0.1, True), nn.Linear(self.dim, self.dim))

    def forward(self, x):
        B, _, H, W = x.shape
        t = NchwToTokensFn.apply(x)
        inter_tok, H1, W1 = self.E_pre.forward_tokens(t, H, W)
        f, H2, W2 = self.E[0].forward_tokens(inter_tok, H1, W1)
        f, H3, W3 = self.E[1].forward_tokens(f, H2, W2)
        fea = TokenMeanFn.apply(f)
        out = linear(linear(fea, self.mlp[0].weight, self.mlp[0].bias, ops.ACT_LRELU, 0.1), self.mlp[2].weight,
                     self.mlp[2].bias)
        inter = TokensToNchwFn.apply(inter_tok, H1, W1)
        inter._fa_tokens = inter_tok
        return fea, [out], inter
