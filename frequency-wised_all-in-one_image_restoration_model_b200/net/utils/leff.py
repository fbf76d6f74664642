"""LeFF parameter container (net/utils/leff.py:71-117).  The math - Linear+GELU, depthwise 3x3 + GELU in
token layout, Linear - runs inside the LeWin block nodes (net/lewin.py: leff_fwd / leff_bwd); the standalone
``forward`` below is the same kernels wrapped for direct use."""
import math

import torch
import torch.nn as nn

from ... import ops


class _LeFFFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, wdw, bdw, w2, b2):
        from ..lewin import leff_fwd
        B, L, C = x.shape
        H = W = int(math.sqrt(L))
        x2 = x.reshape(-1, C).contiguous()
        sv = {}
        y = leff_fwd(x2, w1, b1, wdw, bdw, w2, b2, B, H, W, None, None, sv)
        ctx.geom = (B, H, W)
        ctx.params = (w1, b1, wdw, bdw, w2, b2)
        ctx.save_for_backward(x2, sv['u1'], sv['u2'], sv['h2'], w1, wdw, w2)
        return y.view(B, L, -1)

    @staticmethod
    def backward(ctx, dy):
        from ..lewin import leff_bwd
        x2, u1, u2, h2, w1, wdw, w2 = ctx.saved_tensors
        B, H, W = ctx.geom
        g = dy.reshape(-1, w2.shape[0]).contiguous()
        P_w1, P_b1, P_wdw, P_bdw, P_w2, P_b2 = ctx.params
        dx, (dW1, db1, dwdw, dbdw, dW2, db2) = leff_bwd(g, dict(u1=u1, u2=u2, h2=h2), x2, P_w1, P_b1, P_wdw, P_bdw,
                                                        P_w2, P_b2, B, H, W)
        return dx.view(B, H * W, -1), dW1, db1, dwdw, dbdw, dW2, db2


class LeFF(nn.Module):
    def __init__(self, dim=32, hidden_dim=128, act_layer=nn.GELU, drop=0., use_eca=False, degradation_dim=-1,
                 deform_conv=False):
        super().__init__()
        if deform_conv or use_eca:
            raise NotImplementedError('freqair: the deform_conv / ECA LeFF variants are unreachable at reference HEAD '
                                      '(decoder_Uformer.py:1124 nulls `inter`) and are out of scope')
        if act_layer is not nn.GELU:
            raise NotImplementedError('freqair: LeFF kernels fuse the exact-erf GELU only')
        self.linear1 = nn.Sequential(nn.Linear(dim, hidden_dim), act_layer())
        self.conv = nn.Sequential(nn.Conv2d(hidden_dim, hidden_dim, groups=hidden_dim, kernel_size=3, stride=1,
                                            padding=1), act_layer())
        self.linear2 = nn.Sequential(nn.Linear(hidden_dim, dim))
        self.eca = nn.Identity()

    def forward(self, x, inter=None):
        return _LeFFFn.apply(x, self.linear1[0].weight, self.linear1[0].bias, self.conv[0].weight, self.conv[0].bias,
                             self.linear2[0].weight, self.linear2[0].bias)
