"""Shims that let ``/root/reference/net`` import and run on a CPU-only box (build container only).

Used by ``tools/make_golden*.py`` and ``tests/test_shim_host.py``; never on the GPU box
(``/root/reference`` does not exist there) and never by the product.

* ``timm.models.layers`` (absent): DropPath / to_2tuple / trunc_normal_ restated.  DropPath draws
  its per-sample mask from ``SCALES_RNG`` and appends the scale vector to ``SCALES_LOG`` so a golden
  run can be replayed with explicit scales.
* ``.cuda()`` is neutralised (hard-coded in frequency_decompose.py:17-22, moco.py:161).
* ``DCN_layer.forward`` dies at ``assert False`` (deform_conv.py:64); ``patch_dcn()`` swaps in
  ``torchvision.ops.deform_conv2d`` (stand-in oracle, parity unpinned - SURVEY.md §8c).
"""
import sys
import types

import torch
import torch.nn as nn

SCALES_LOG = []
SCALES_RNG = torch.Generator()


def install(argv=()):
    if 'timm.models.layers' in sys.modules and getattr(sys.modules['timm.models.layers'], '_freqair_shim', False):
        return
    sys.argv = ['ref'] + list(argv)

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            r = torch.empty(x.shape[0]).bernoulli_(keep, generator=SCALES_RNG) / keep
            SCALES_LOG.append(r.clone())
            return x * r.view(-1, *([1] * (x.ndim - 1)))

    def to_2tuple(x):
        return x if isinstance(x, tuple) else (x, x)

    layers = types.ModuleType('timm.models.layers')
    layers.DropPath, layers.to_2tuple, layers.trunc_normal_ = DropPath, to_2tuple, nn.init.trunc_normal_
    layers._freqair_shim = True
    sys.modules['timm'] = types.ModuleType('timm')
    sys.modules['timm.models'] = types.ModuleType('timm.models')
    sys.modules['timm.models.layers'] = layers
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    if '/root/reference' not in sys.path:
        sys.path.insert(0, '/root/reference')


def patch_dcn():
    from torchvision.ops import deform_conv2d
    from net.utils import deform_conv as dc

    def forward(self, input_feat, inter):
        out = self.conv_offset_mask(torch.cat([input_feat, inter], dim=1))
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        return deform_conv2d(input_feat, torch.cat((o1, o2), dim=1), self.weight, self.bias,
                             stride=self.stride, padding=self.padding, dilation=self.dilation,
                             mask=torch.sigmoid(mask))
    dc.DCN_layer.forward = forward
