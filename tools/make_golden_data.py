"""Generate ``tests/golden/datagen.npz`` from the UNMODIFIED reference data-pipeline functions (build container only).

Run:  python tools/make_golden_data.py        (needs /root/reference)

Calls utils/image_utils.data_augmentation, crop_img and utils/dataset_utils._crop_patch exactly as
TrainDataset.__getitem__ does (dataset_utils.py:118-135) on a small seeded uint8 image, with Python's ``random`` seeded
so the crop origins are reproducible, and stores the inputs, the draws and the reference's outputs.
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, '/root/reference')
from torchvision.transforms import ToTensor                      # noqa: E402
from utils.dataset_utils import _crop_patch                       # noqa: E402
from utils.image_utils import crop_img, data_augmentation         # noqa: E402

rng = np.random.RandomState(7)
raw = rng.randint(0, 256, size=(45, 38, 3)).astype(np.uint8)
gt = crop_img(raw, base=16)                                       # dataset_utils.py:118
noise = rng.randn(*gt.shape)
sigma = 25
noisy = np.clip(gt + noise * sigma, 0, 255).astype(np.uint8)      # dataset_utils.py:126
P = 12
out = {'raw': raw, 'gt': gt, 'noise': noise.astype(np.float32), 'sigma': np.array(sigma), 'noisy': noisy, 'P': np.array(P)}
tt = ToTensor()
random.seed(11)
for mode in range(8):
    state = random.getstate()
    ind_h = random.randint(0, gt.shape[0] - P)
    ind_w = random.randint(0, gt.shape[1] - P)
    random.setstate(state)
    d, c = _crop_patch(noisy, gt, size=P)                          # dataset_utils.py:130
    import torch
    if mode == 0:
        d_aug, c_aug = data_augmentation(torch.from_numpy(d), 0), data_augmentation(torch.from_numpy(c), 0)
    else:
        d_aug, c_aug = data_augmentation(d, mode).copy(), data_augmentation(c, mode).copy()
    out[f'origin{mode}'] = np.array([ind_h, ind_w])
    out[f'deg{mode}'] = tt(np.ascontiguousarray(d_aug)).numpy()
    out[f'clean{mode}'] = tt(np.ascontiguousarray(c_aug)).numpy()
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'datagen.npz'), **out)
print('wrote tests/golden/datagen.npz', {k: v.shape for k, v in out.items() if k.startswith('deg')})
