"""Loss nodes used by the train step (train.py:64-65,88-92)."""
import torch

from . import ops


class L1LossFn(torch.autograd.Function):
    """nn.L1Loss() (mean absolute error) - forward and the sign gradient in one streaming kernel."""

    @staticmethod
    def forward(ctx, a, b):
        loss, grad = ops.l1_loss(a.contiguous(), b.contiguous(), want_grad=True)
        ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def l1_loss(a, b):
    return L1LossFn.apply(a, b)


class SpectralL1Fn(torch.autograd.Function):
    """l1(decompose(a), decompose(b)) with decompose = FrequencyDecompose('frequency_decompose', 1/nb, n, n,
    inverse=False) (train.py:69-70,90-91): one fused FFT pass per map, the band stack is never materialised."""

    @staticmethod
    def forward(ctx, a, b, bob, nbands):
        loss, grad = ops.spectral_l1(a.contiguous(), b.contiguous(), bob, nbands, want_grad=True)
        ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


def spectral_l1_loss(a, b, decompose):
    """``decompose``: the net.utils.frequency_decompose.FrequencyDecompose(..., inverse=False) instance of train.py:70."""
    assert decompose.type == 'frequency_decompose' and decompose.inverse is False
    return SpectralL1Fn.apply(a, b, decompose.band_of_bin(a.device), decompose.out_bands)
