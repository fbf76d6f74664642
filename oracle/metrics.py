"""PSNR / SSIM as the reference's evaluation computes them (utils/val_utils.py:50-66 calls
skimage.metrics.peak_signal_noise_ratio / structural_similarity with data_range=1, channel_axis=2 on images clipped
to [0,1]).  skimage is absent from this image, so its published definitions are restated here (test infrastructure):
  PSNR = 10 log10(data_range^2 / MSE)
  SSIM (Wang et al. 2004, skimage defaults): 7x7 uniform window, K1=0.01, K2=0.03, sample covariance (N/(N-1)),
  mean over the window-valid interior, averaged over channels."""
import torch
import torch.nn.functional as F


def psnr(a, b, data_range=1.0):
    a, b = a.double().clamp(0, 1), b.double().clamp(0, 1)
    mse = (a - b).pow(2).mean()
    return float(10.0 * torch.log10(data_range ** 2 / mse))


def ssim(a, b, data_range=1.0, win=7):
    """a, b: [C,H,W] (or [B,C,H,W]) in [0,1]."""
    a, b = a.double().clamp(0, 1), b.double().clamp(0, 1)
    if a.dim() == 3:
        a, b = a[None], b[None]
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    NP = win * win
    cov_norm = NP / (NP - 1.0)
    pool = lambda t: F.avg_pool2d(t, win, stride=1)                    # == uniform_filter cropped to the valid interior
    ux, uy = pool(a), pool(b)
    vx = cov_norm * (pool(a * a) - ux * ux)
    vy = cov_norm * (pool(b * b) - uy * uy)
    vxy = cov_norm * (pool(a * b) - ux * uy)
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    return float(S.mean())
