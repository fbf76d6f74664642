"""Oracle: the reference's tiled full-resolution inference (test.py:43-71), test infrastructure (see oracle/__init__.py).

One deviation, documented wherever it matters: test.py:67 accumulates ``patched_input_img`` - the
degraded tiles - so the image it reassembles is its own input.  This restatement, like the product (infer.py) and the
eval-set golden (tools/make_golden_eval.py), accumulates the RESTORED tiles (``patched_restored``)."""
import torch


def tile_origins(H, W, patch=128):
    """test.py:48-49: regular grid, last row / column anchored at H - patch / W - patch (overlapping)."""
    assert H >= patch and W >= patch and patch % 8 == 0                        # test.py:43-45
    return list(range(0, H - patch, patch)) + [H - patch], list(range(0, W - patch, patch)) + [W - patch]


def restore_tiled(forward, img, patch=128):
    """img [1,C,H,W]; ``forward(tiles [T,C,patch,patch]) -> restored tiles`` (the eval forward, test.py:59)."""
    assert img.shape[0] == 1
    _, C, H, W = img.shape
    hs, ws = tile_origins(H, W, patch)
    tiles = torch.cat([img[..., h:h + patch, w:w + patch] for h in hs for w in ws], 0)     # test.py:51-57
    restored = forward(tiles)
    E, Wt = torch.zeros(C, H, W, dtype=img.dtype), torch.zeros(C, H, W, dtype=img.dtype)
    cnt = 0
    for h in hs:                                                               # test.py:61-69
        for w in ws:
            E[..., h:h + patch, w:w + patch].add_(restored[cnt])
            Wt[..., h:h + patch, w:w + patch].add_(1.0)
            cnt += 1
    return (E / Wt).unsqueeze(0)
