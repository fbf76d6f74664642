"""Oracle: frequency-band decomposition (test infrastructure; see oracle/__init__.py).

Restates ``net/utils/frequency_decompose.py`` of the reference.  The reference
does ``fft2 -> fftshift -> radial mask -> ifftshift -> ifft2().real`` per band;
shifting the *mask* instead of the spectrum is the same thing, so the oracle
builds un-shifted masks once and applies them to ``torch.fft.fft2`` directly.
"""
import math

import torch


def band_index_map(kind: str, size: float, h: int, w: int) -> torch.Tensor:
    """int64 [h, w] map: band id of every *un-shifted* FFT bin.

    kind 'frequency_decompose'   (frequency_decompose.py:28-68): bands
        [s_{i-1} R, s_i R) for s in linspace(size, 1, nb), the last one closed.
    kind 'frequency_decompose_1' (frequency_decompose.py:70-107): nb+1 bands,
        dist <= s_i R for s in linspace(0, 1, nb+1) XOR the previous disc, so
        band 0 is the DC bin alone.
    Geometry follows frequency_decompose.py:17-26: integer grids, centre
    (int(w/2), int(h/2)), fp32 sqrt, R = sqrt(cx^2 + cy^2).
    """
    Y = torch.arange(h).unsqueeze(1)
    X = torch.arange(w).unsqueeze(0)
    cx, cy = int(w / 2), int(h / 2)
    dist = torch.sqrt(((X - cx) ** 2 + (Y - cy) ** 2).to(torch.float32))
    max_radius = torch.sqrt(torch.tensor(cx ** 2 + cy ** 2, dtype=torch.float32))
    nb = math.floor(1.0 / size + 0.1)
    last = torch.zeros(h, w, dtype=torch.bool)
    idx = torch.full((h, w), -1, dtype=torch.int64)
    if kind == 'frequency_decompose':
        steps = torch.linspace(size, 1, nb)
    elif kind == 'frequency_decompose_1':
        steps = torch.linspace(0, 1, nb + 1)
    else:
        raise ValueError(kind)
    for i, sz in enumerate(steps):
        radius = max_radius * sz
        if kind == 'frequency_decompose' and not sz == 1.0:
            mask = dist < radius
        else:
            mask = dist <= radius
        now = mask ^ last
        last = mask
        idx[now] = i
    # masks are defined on the fftshift-ed spectrum; bring them back to FFT order
    return torch.fft.ifftshift(idx)


def num_bands(kind: str, size: float) -> int:
    nb = math.floor(1.0 / size + 0.1)
    return {'frequency_decompose': nb, 'frequency_decompose_1': nb + 1}.get(kind, 2)


def decompose(x: torch.Tensor, kind: str, size: float, inverse=True) -> torch.Tensor:
    """FrequencyDecompose(kind, size, h, w, inverse)(x) -> [bands, ...]  (frequency_decompose.py:120-125)."""
    if kind not in ('frequency_decompose', 'frequency_decompose_1'):
        # frequency_decompose_dc (frequency_decompose.py:109-118): mean / residual
        d = x.mean(-1, keepdim=True).mean(-2, keepdim=True).expand_as(x)
        return torch.stack([d, x - d], 0)
    h, w = x.shape[-2:]
    idx = band_index_map(kind, size, h, w)
    spec = torch.fft.fft2(x)
    outs = []
    for i in range(num_bands(kind, size)):
        part = spec * (idx == i)
        if inverse == 'visual':                               # :53-54 abs of the shifted spectrum; the
            outs.append(torch.fft.fftshift(part.abs()))       # reference's dim-less fftshift (:32) also rolls B and C
        elif inverse is True:                                 # :55-57
            outs.append(torch.fft.ifft2(part).real)
        elif inverse is False:                                # :58-60
            outs.append(torch.stack((part.real, part.imag), -1))
        else:
            raise ValueError(inverse)
    return torch.stack(outs, 0)


def band_filter(x: torch.Tensor, kind: str, size: float, coef: torch.Tensor) -> torch.Tensor:
    """x + sum_i coef[i] * band_i(x): the only way bands are consumed on the hot path
    (decoder_Uformer.py:275-288 with coef[0]=0; encoder_ViT.py:85-92).

    coef: [bands, *broadcastable to x.shape[:-2]].
    """
    bands = decompose(x, kind, size, True)
    return x + (bands * coef[..., None, None]).sum(0)
