"""Deterministic synthetic inputs of the benchmark / parity shapes.

No dataset ships with the reference (``data/`` is absent), so every measured or
parity-checked run uses tensors that are a pure function of a seed.  The
degradation recipes mirror the reference's data pipeline where it has one:
Gaussian noise is added in 0..255 space, clipped and quantised to uint8 before
``ToTensor`` (``utils/dataset_utils.py:122-126``).  Host-side torch only.
"""
import torch
import torch.nn.functional as F


def clean_images(B, H=128, W=128, seed=1234):
    """Box-blurred uniform noise in [0,1] (decaying spectrum so every radial band is populated)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, H + 4, W + 4, generator=g)
    x = F.avg_pool2d(x, 5, stride=1)
    return x.clamp_(0, 1).contiguous()


def quantise(x255):
    return (x255.clamp(0, 255).to(torch.uint8).to(torch.float32) / 255.0)


def gaussian_noise(clean, sigma, seed=1235):
    g = torch.Generator().manual_seed(seed)
    n = torch.randn(clean.shape, generator=g)
    return quantise(clean * 255.0 + n * sigma)


def rain(clean, seed=1236, density=0.01, strength=0.6):
    """Sparse 45-degree streaks (1x9) added to the clean image."""
    g = torch.Generator().manual_seed(seed)
    B, C, H, W = clean.shape
    seeds = (torch.rand(B, 1, H, W, generator=g) < density).float()
    k = torch.eye(9).flip(1).view(1, 1, 9, 9)
    streak = F.conv2d(seeds, k, padding=4).clamp_(0, 1)
    return quantise((clean + strength * streak) * 255.0)


def haze(clean, t=0.6, A=0.8):
    return quantise((clean * t + A * (1 - t)) * 255.0)


DEGRADATIONS = ('sigma15', 'sigma25', 'sigma50', 'rain', 'haze')


def degrade(clean, kind, seed=1235):
    if kind.startswith('sigma'):
        return gaussian_noise(clean, float(kind[5:]), seed)
    if kind == 'rain':
        return rain(clean, seed)
    if kind == 'haze':
        return haze(clean)
    raise ValueError(kind)


def mixed_batch(B, H=128, W=128, seed=1234, kinds=DEGRADATIONS):
    """(x_query, x_key, clean): sample i cycles through ``kinds`` (BASELINE config 3);
    x_key is a second, independent degradation draw of the same clean crop."""
    clean = clean_images(B, H, W, seed)
    xq = torch.empty_like(clean)
    xk = torch.empty_like(clean)
    for i in range(B):
        kind = kinds[i % len(kinds)]
        xq[i:i + 1] = degrade(clean[i:i + 1], kind, seed + 1 + 7 * i)
        xk[i:i + 1] = degrade(clean[i:i + 1], kind, seed + 2 + 7 * i)
    return xq, xk, clean


def noisy_batch(B, sigma=25, H=128, W=128, seed=1234):
    """(x_query, x_key, clean) with Gaussian noise only (BASELINE configs 1, 2)."""
    clean = clean_images(B, H, W, seed)
    return gaussian_noise(clean, sigma, seed + 1), gaussian_noise(clean, sigma, seed + 2), clean


def tile_indices(H, W, patch=128):
    """Tile origins of the reference's tiled inference (test.py:48-49)."""
    hs = list(range(0, H - patch, patch)) + [H - patch]
    ws = list(range(0, W - patch, patch)) + [W - patch]
    return hs, ws
