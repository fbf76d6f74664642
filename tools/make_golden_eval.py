"""PSNR / SSIM of the UNMODIFIED reference network on the synthetic eval set of SURVEY.md section 8(d) (build container
only; needs /root/reference, ~6 min CPU):

  32 synthetic 256 x 256 images (seed 4321), sigma = 25, tiled by test.py:48-57 into 128 x 128 patches, one batched eval
  forward per image (test.py:59), overlap-averaged reassembly of the RESTORED tiles - the reference accumulates
  ``patched_input_img`` (test.py:67), i.e. its reassembled image is the input; that defect is the one documented
  deviation - name-keyed deterministic weights (oracle.detfill), PSNR / SSIM as utils/val_utils.py:50-66 defines them
  (oracle.metrics restates skimage, absent here).

Writes tests/golden/evalset_uu.npz: per-image psnr / ssim of the reference's restored image against the clean image, the
first restored image in full, and a strided sample of every restored image."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import ref_shims  # noqa: E402
from make_golden import strided_sample  # noqa: E402
from oracle import detfill, metrics  # noqa: E402

synth = importlib.import_module('frequency-wised_all-in-one_image_restoration_model_b200.synth')
N_IMG, SIZE, PATCH = 32, 256, 128


def eval_images():
    clean = synth.clean_images(N_IMG, SIZE, SIZE, seed=4321)
    noisy = torch.cat([synth.gaussian_noise(clean[i:i + 1], 25, 4322 + i) for i in range(N_IMG)])
    return noisy, clean


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref_shims.install(['--degradation_embedding_method', 'all_3_bands'])
    from net.model import AirNet
    from option import options as opt
    opt.batch_size = 2
    net = AirNet(opt)
    detfill.fill_state(net)
    net.eval()
    noisy, clean = eval_images()
    psnr, ssim, samp, first = [], [], [], None
    for i in range(N_IMG):
        img = noisy[i:i + 1]
        _, C, H, W = img.shape
        hs = list(range(0, H - PATCH, PATCH)) + [H - PATCH]                 # test.py:48-49
        ws = list(range(0, W - PATCH, PATCH)) + [W - PATCH]
        tiles = torch.cat([img[..., h:h + PATCH, w:w + PATCH] for h in hs for w in ws], 0)
        with torch.no_grad():
            restored_tiles = net(x_query=tiles, x_key=tiles)               # test.py:59
        E, Wt = torch.zeros(C, H, W), torch.zeros(C, H, W)
        cnt = 0
        for h in hs:
            for w in ws:
                E[..., h:h + PATCH, w:w + PATCH].add_(restored_tiles[cnt])  # test.py:67 with patched_restored
                Wt[..., h:h + PATCH, w:w + PATCH].add_(1.0)
                cnt += 1
        restored = E / Wt
        psnr.append(metrics.psnr(restored, clean[i]))
        ssim.append(metrics.ssim(restored, clean[i]))
        samp.append(strided_sample(restored, 4096).numpy().copy())
        if first is None:
            first = restored.numpy().copy()
        print(i, psnr[-1], ssim[-1], flush=True)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'evalset_uu.npz'), psnr=np.array(psnr), ssim=np.array(ssim),
                        samp=np.stack(samp), first=first)


if __name__ == '__main__':
    main()
