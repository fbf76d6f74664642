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


# ----------------------------------------------------------------------------- part 2: the dataset object itself
def dataset_fixture():
    """Run the UNMODIFIED TrainDataset (utils/dataset_utils.py:68-143) on a tiny PNG tree with a seeded ``random`` and
    store what it returns, so that the host-side draw logic of ``datagen.DeviceTrainSet`` (type round-robin, shuffle at
    wrap-around, two crop + augmentation draws per item) is pinned to the real thing.  Paired types only (the noise
    synthesis of 'denoising_*' draws np.random.randn, which DeviceTrainSet replaces by its stateless generator)."""
    import tempfile
    import types
    from PIL import Image
    from utils.dataset_utils import TrainDataset
    rng2 = np.random.RandomState(21)
    cwd = os.getcwd()
    out2 = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            spec = {'deraining': [(40, 52), (37, 45), (33, 64)], 'dehazing': [(48, 48), (36, 70)]}
            for de, sizes in spec.items():
                os.makedirs(f'data/{de}_train/GT'); os.makedirs(f'data/{de}_train/Input')
                for i, (h, w) in enumerate(sizes):
                    gt_img = rng2.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
                    in_img = rng2.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
                    Image.fromarray(gt_img).save(f'data/{de}_train/GT/{de[:3]}{i}.png')
                    Image.fromarray(in_img).save(f'data/{de}_train/Input/{de[:3]}{i}_x.png')
                    out2[f'img/{de}/{de[:3]}{i}/gt'] = gt_img
                    out2[f'img/{de}/{de[:3]}{i}/in'] = in_img
            args = types.SimpleNamespace(de_type=['deraining', 'dehazing'], patch_size=16)
            ds = TrainDataset(args)
            for t, de in enumerate(args.de_type):             # listdir order is file-system dependent: record it
                out2[f'order/{de}'] = np.array([os.path.basename(p).split('.')[0] for p in ds.gt_ids[t]])
            random.seed(5)
            n = 9                                              # wraps both per-type iterators at least once
            names, des = [], []
            for k in range(n):
                (name, de), d1, d2, c1, c2 = ds[k]
                names.append(name); des.append(de)
                out2[f'item{k}/d1'], out2[f'item{k}/d2'] = d1.numpy(), d2.numpy()
                out2[f'item{k}/c1'], out2[f'item{k}/c2'] = c1.numpy(), c2.numpy()
            out2['names'] = np.array(names); out2['de_ids'] = np.array(des); out2['seed'] = np.array(5); out2['n'] = np.array(n)
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'datagen_dataset.npz'), **out2)
    print('wrote tests/golden/datagen_dataset.npz', names, des)


dataset_fixture()
