"""DCN_layer - parameter container with the reference's names and initialisation
(net/utils/deform_conv.py:11-54).  The reference's forward stops at ``assert False`` (:64) because its mmcv
call is commented out; here the DCNv2 arithmetic (K7: NHWC bilinear gather + dense contraction) runs inside
the DGM node of net/decoder_DGRN.py.  Semantics are pinned to torchvision.ops.deform_conv2d's DCNv2
(**parity unpinned** upstream - see oracle/airnet.py)."""
import math

import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair


class DCN_layer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True, extra_offset_mask=True):
        super().__init__()
        assert _pair(kernel_size) == (3, 3) and stride == 1 and padding == 1 and dilation == 1 and groups == 1 \
            and deformable_groups == 1, 'freqair: DCNv2 kernels cover the 3x3 / s1 / p1 / single-group case DGRN uses'
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deformable_groups, self.with_bias = groups, deformable_groups, bias
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        self.extra_offset_mask = extra_offset_mask
        self.conv_offset_mask = nn.Conv2d(self.in_channels * 2, self.deformable_groups * 3 * 9,
                                          kernel_size=self.kernel_size, stride=_pair(stride), padding=_pair(padding),
                                          bias=True)
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.init_offset()
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def init_offset(self):
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()
