"""CPU: the host-side draw logic of datagen.DeviceTrainSet against the UNMODIFIED reference TrainDataset.

``tests/golden/datagen_dataset.npz`` (tools/make_golden_data.py) holds what utils/dataset_utils.TrainDataset returned for
nine consecutive items on a tiny PNG tree with ``random.seed(5)``.  Here the same images and the same seeded generator
go through DeviceTrainSet; the one CUDA call it makes (ops.crop_augment) is replaced by the oracle's numpy restatement -
this test checks WHICH image, crop origin and augmentation are drawn and in what order, the kernel itself is checked on
the GPU (tests/test_gpu_ops.py)."""
import importlib
import os
import random

import numpy as np
import torch

from conftest import PKG_NAME
from oracle import datagen as od


def oracle_crop_augment(pool, meta, sigma, noise, P):
    pool = pool.numpy()
    deg, clean = [], []
    for m in meta.tolist():
        c_off, d_off, H, W, y0, x0, mode, _ = m
        assert d_off >= 0, 'paired types only in this fixture'
        cimg = pool[c_off:c_off + H * W * 3].reshape(H, W, 3)
        dimg = pool[d_off:d_off + H * W * 3].reshape(H, W, 3)
        d, c = od.training_pair(cimg, dimg, y0, x0, mode, P)
        deg.append(d); clean.append(c)
    return torch.from_numpy(np.stack(deg)), torch.from_numpy(np.stack(clean))


def test_device_train_set_replays_reference_dataset(monkeypatch):
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'datagen_dataset.npz'))
    dg = importlib.import_module(PKG_NAME + '.datagen')
    ops = importlib.import_module(PKG_NAME + '.ops')
    monkeypatch.setattr(ops, 'crop_augment', oracle_crop_augment)
    de_type = ['deraining', 'dehazing']
    images = {de: [(str(name), g[f'img/{de}/{name}/gt'], g[f'img/{de}/{name}/in']) for name in g[f'order/{de}']]
              for de in de_type}
    ds = dg.DeviceTrainSet(de_type, images, patch_size=16, device='cpu', rng=random.Random(int(g['seed'])))
    n = int(g['n'])
    got = []
    for B in (4, 5):                                   # two batches: the iterators carry over between calls
        (names, de_ids), d1, d2, c1, c2 = ds.next_batch(B)
        for i in range(B):
            got.append((names[i], de_ids[i], d1[i].numpy(), d2[i].numpy(), c1[i].numpy(), c2[i].numpy()))
    assert len(got) == n
    for k, (name, de, d1, d2, c1, c2) in enumerate(got):
        assert name == str(g['names'][k]) and de == str(g['de_ids'][k]), (k, name, de)
        for mine, key in ((d1, 'd1'), (d2, 'd2'), (c1, 'c1'), (c2, 'c2')):
            assert np.array_equal(mine, g[f'item{k}/{key}']), (k, key)
    assert len(ds) == 400 * len(de_type)
