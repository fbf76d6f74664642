"""Tiled full-resolution inference, the reference's test.py:48-71 on the device: an image is cut into 128x128 tiles
(the last row / column anchored at H-128 / W-128, overlapping), every tile goes through ONE batched eval forward, and
the restored tiles are overlap-averaged back.

Deviation, on purpose and documented: the reference accumulates ``patched_input_img`` - the
degraded tiles - instead of the network output (test.py:67), so its reassembled image is the input.  This front end
reassembles the RESTORED tiles."""
import torch

from .synth import tile_indices


def tile(img, patch=128):
    """[1,3,H,W] -> ([T,3,patch,patch], origins)"""
    _, _, H, W = img.shape
    assert H >= patch and W >= patch and patch % 8 == 0        # test.py:43-45
    hs, ws = tile_indices(H, W, patch)
    tiles = torch.stack([img[0, :, h:h + patch, w:w + patch] for h in hs for w in ws])
    return tiles.contiguous(), [(h, w) for h in hs for w in ws]


def untile(tiles, origins, H, W, patch=128):
    out = torch.zeros(1, tiles.shape[1], H, W, device=tiles.device, dtype=tiles.dtype)
    cnt = torch.zeros(1, 1, H, W, device=tiles.device, dtype=tiles.dtype)
    for t, (h, w) in zip(tiles, origins):
        out[0, :, h:h + patch, w:w + patch] += t
        cnt[0, :, h:h + patch, w:w + patch] += 1
    return out / cnt


@torch.no_grad()
def restore_tiled(net, img, patch=128):
    """net: AirNet in eval mode; img [1,3,H,W] on the device -> restored [1,3,H,W]."""
    tiles, origins = tile(img, patch)
    restored = net(x_query=tiles, x_key=tiles)                   # test.py:59
    return untile(restored, origins, img.shape[2], img.shape[3], patch)


class GraphedRestorer:
    """restore_tiled for one image size, captured in a CUDA graph (a 512x512 image is ~560 library launches plus the
    torch glue of tiling: launch-bound when issued one by one).  ``__call__(img)`` copies the image into the static
    input buffer, replays, and returns the static output tensor (valid until the next call)."""

    def __init__(self, net, H, W, patch=128, warmup=2):
        self.net, self.patch = net, patch
        dev = next(net.parameters()).device
        self.static_in = torch.zeros(1, 3, H, W, device=dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                restore_tiled(net, self.static_in, patch)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import ops
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_out = restore_tiled(net, self.static_in, patch)
        self.graph_launches = ops.launch_count() - n0        # libfreqair kernels recorded in the graph (per replay)

    def __call__(self, img):
        self.static_in.copy_(img, non_blocking=True)
        self.graph.replay()
        return self.static_out
