"""Deterministic, name-keyed parameter filling (test infrastructure).

The reference's weights for a parity run cannot travel (373 M parameters), and
identical-seed construction is fragile across implementations, so every parity
artefact is produced from weights that are a pure function of
``(parameter name, shape)``: the reference run in the build container
(``tools/make_golden.py``), the oracle and the CUDA modules on the GPU box all
call :func:`fill_state` and get bit-identical tensors.
"""
import zlib

import torch


def _gen(name: str) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return g


def det_tensor(name: str, shape, like: torch.Tensor = None) -> torch.Tensor:
    """Value for one state_dict entry.  Integer buffers are returned as ``like``."""
    if like is not None and not like.is_floating_point():
        if name.endswith('num_batches_tracked') or name.endswith('queue_ptr'):
            return torch.zeros_like(like)
        return like.clone()
    shape = tuple(shape)
    if name.split('.')[-1] == 'mask_freq':    # structural 0/-100 constant (encoder_Uformer.py:246-254): keep
        return like.clone() if like is not None else torch.zeros(shape)
    g = _gen(name)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    leaf = name.split('.')[-1]
    if leaf == 'running_var':
        return 1.0 + 0.2 * r.abs()
    if leaf == 'running_mean':
        return 0.1 * r
    if leaf == 'queue':                       # MoCo queue: unit columns (moco.py:38-40)
        return torch.nn.functional.normalize(r, dim=1)
    if 'relative_position_bias_table' in name:
        return 0.3 * r
    if leaf == 'lamb':                        # ViT band weights (encoder_ViT.py:63)
        return 0.5 * r
    if leaf == 'pos_embedding':
        return 0.2 * r
    if len(shape) == 1:
        # affine scale of a norm layer vs. plain bias
        if leaf == 'weight':
            return 1.0 + 0.1 * r
        return 0.05 * r
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    gain = 0.3 if 'conv_offset_mask' in name else 0.7
    return r * (gain / max(fan_in, 1) ** 0.5)


def fill_state(module_or_sd):
    """Overwrite every entry of a module's state_dict (or a dict) in place; returns the dict."""
    sd = module_or_sd if isinstance(module_or_sd, dict) else module_or_sd.state_dict()
    with torch.no_grad():
        for k, v in sd.items():
            v.copy_(det_tensor(k, v.shape, v).to(v.dtype))
    return sd


def make_state(spec):
    """Build a fresh CPU state dict from ``{name: (shape, dtype_str)}``."""
    out = {}
    for k, (shape, dt) in spec.items():
        dtype = getattr(torch, dt)
        proto = torch.zeros(tuple(shape), dtype=dtype)
        out[k] = det_tensor(k, shape, proto).to(dtype)
    return out
