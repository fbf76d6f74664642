"""``FrequencyDecompose(type, size, h, w, inverse=True)`` - same constructor and call contract as the reference
(net/utils/frequency_decompose.py:5-125; imported directly by train.py:16,70), arithmetic in the K1 kernels.

The radial band geometry is rebuilt on the host exactly as the reference does it (integer grids, fp32 sqrt,
``torch.linspace`` band edges, :17-26,38-48,80-88) and handed to the kernels as a uint8 band-id map of the
un-shifted half spectrum, so fftshift / ifftshift / mask.repeat never run on the device.
"""
import math

import torch
from torch import nn

from ... import ops


def band_index_map(kind, size, h, w):
    """int64 [h, w]: band id of every un-shifted FFT bin (-1 = in no band)."""
    Y = torch.arange(h).unsqueeze(1)
    X = torch.arange(w).unsqueeze(0)
    cx, cy = int(w / 2), int(h / 2)
    dist = torch.sqrt(((X - cx) ** 2 + (Y - cy) ** 2).to(torch.float32))
    max_radius = torch.sqrt(torch.tensor(cx ** 2 + cy ** 2, dtype=torch.float32))
    nb = math.floor(1. / size + 0.1)
    edges = torch.linspace(size, 1, nb) if kind == 'frequency_decompose' else torch.linspace(0, 1, nb + 1)
    idx = torch.full((h, w), -1, dtype=torch.int64)
    inside_prev = torch.zeros(h, w, dtype=torch.bool)
    for i, sz in enumerate(edges):
        r = max_radius * sz
        inside = (dist < r) if (kind == 'frequency_decompose' and not sz == 1.0) else (dist <= r)
        idx[inside ^ inside_prev] = i
        inside_prev = inside
    return torch.fft.ifftshift(idx)


def half_band_map(kind, size, n):
    """uint8 [n, n/2+1] band ids of the rfft2 half spectrum (the kernels' ``band_of_bin``)."""
    m = band_index_map(kind, size, n, n)
    assert int(m.min()) >= 0, 'every bin must belong to a band'
    return m[:, :n // 2 + 1].to(torch.uint8).contiguous()


class _BandSplit(torch.autograd.Function):
    """y[band] = irfft2(rfft2(x) * mask_band); each band projector is self-adjoint."""

    @staticmethod
    def forward(ctx, x, bob, nbands):
        ctx.bob, ctx.nbands = bob, nbands
        return ops.band_split(x.contiguous(), bob, nbands, 0)

    @staticmethod
    def backward(ctx, dy):
        # sum_b P_b(dy_b): filter every band's gradient with its own mask, then add
        dy = dy.contiguous()
        n = dy.shape[-1]
        out = None
        for b in range(ctx.nbands):
            coef = torch.zeros(1, ctx.nbands, device=dy.device)
            coef[0, b] = 1.0
            part = ops.band_filter(dy[b], ctx.bob, coef, dy[b].numel() // (n * n), 1) - dy[b]
            out = part if out is None else out + part
        return out, None, None


class FrequencyDecompose(nn.Module):
    def __init__(self, type, size, h, w, inverse=True):
        super().__init__()
        self.type, self.size, self.h, self.w, self.inverse = type, size, h, w, inverse
        assert size > 0 and size <= 1, 'invalid frequency band width(size=%s)' % (size)
        self._bob = None
        if self.type in ['frequency_decompose', 'frequency_decompose_1']:
            assert h == w and (h & (h - 1)) == 0 and 8 <= h <= 128, 'freqair: square power-of-two maps in 8..128 only'
            nb = math.floor(1. / size + 0.1)
            self.num_bands = nb
            self.out_bands = nb if self.type == 'frequency_decompose' else nb + 1
            self._bob_cpu = half_band_map(self.type, size, h)

    def band_of_bin(self, device):
        if self._bob is None or self._bob.device != device:
            self._bob = self._bob_cpu.to(device)
        return self._bob

    def forward(self, x):
        if self.type in ['frequency_decompose', 'frequency_decompose_1']:
            bob = self.band_of_bin(x.device)
            if self.inverse is True:
                if x.requires_grad and torch.is_grad_enabled():
                    return _BandSplit.apply(x, bob, self.out_bands)
                return ops.band_split(x.contiguous(), bob, self.out_bands, 0)
            if self.inverse is False:
                if x.requires_grad and torch.is_grad_enabled():
                    raise RuntimeError('freqair: spectrum output (inverse=False) has no backward yet')
                return ops.band_split(x.contiguous(), bob, self.out_bands, 1)
            raise RuntimeError("freqair: inverse='visual' (debug plots) is outside the accelerated path")
        return ops.dc_split(x.contiguous())
