"""Device-resident training-pair generator: ``TrainDataset`` + ``DataLoader(batch_size=B)`` of the reference
(utils/dataset_utils.py:68-143, train.py:47-52) with the images held in HBM and the pixel work done by ONE kernel
launch per batch (``fa_crop_augment``), instead of 16 CPU workers decoding, cropping and augmenting one pair at a time.

What stays on the host, drawn exactly as the reference draws it (same ``random`` calls in the same order, so a seeded
``random`` gives the same image order, crop origins and augmentation ids): the round-robin over degradation types
(:98,138), the per-type iterator and its shuffle at wrap-around (:100-104), ``np.random.choice([15, 25, 50])`` for
``denoising_0`` (:123-125), two ``(ind_H, ind_W, flag_aug)`` draws per sample (:130-131; image_utils.py:178-182).
What differs, by design: the Gaussian noise field comes from a stateless per-pixel generator (csrc/datagen.cu) rather
than ``np.random.randn`` over the whole image - same distribution, same sharing of noise between the two overlapping
crops of a pair, different stream.
"""
import random

import numpy as np
import torch

from . import ops


def crop_img(image, base=16):
    """utils/image_utils.crop_img (:59-64): centre-crop H and W to multiples of ``base``."""
    h, w = image.shape[0], image.shape[1]
    ch, cw = h % base, w % base
    return image[ch // 2:h - ch + ch // 2, cw // 2:w - cw + cw // 2, :]


class DeviceTrainSet:
    """``de_type``: list of degradation names as in ``opt.de_type`` ('denoising_15', 'denoising_0' (random sigma),
    'deraining', 'dehazing', ...).  ``images[name]``: list of ``(gt_name, clean HWC uint8, degraded HWC uint8 or None)``;
    ``None`` is only valid for 'denoising_*' types, whose degradation is synthesised."""

    def __init__(self, de_type, images, patch_size=128, device='cuda', rng=random, np_rng=np.random, seed=0):
        self.de_type = list(de_type)
        self.patch_size = patch_size
        self.device = torch.device(device)
        self.rng, self.np_rng = rng, np_rng
        self.de_type_iterator = 0
        self.de_iterator = [0] * len(self.de_type)
        self.seed = seed
        self.entries = []                      # per type: list of dicts (order = the reference's gt_ids / input_ids order)
        chunks, off = [], 0
        for name in self.de_type:
            ents = []
            for gt_name, clean, degraded in images[name]:
                clean = np.ascontiguousarray(crop_img(np.asarray(clean, np.uint8), 16))           # :118
                H, W = clean.shape[:2]
                assert H >= patch_size and W >= patch_size, 'image smaller than the patch'
                e = dict(name=gt_name, H=H, W=W, clean=off, degraded=-1, uid=len(chunks))
                chunks.append(clean.reshape(-1)); off += clean.size
                if degraded is not None:
                    degraded = np.ascontiguousarray(crop_img(np.asarray(degraded, np.uint8), 16))   # :128
                    assert degraded.shape == clean.shape
                    e['degraded'] = off
                    chunks.append(degraded.reshape(-1)); off += degraded.size
                else:
                    assert 'denoising' in name, 'only denoising types synthesise their degradation'
                ents.append(e)
            self.entries.append(ents)
        self.pool = torch.from_numpy(np.concatenate(chunks)).to(self.device)
        self.epoch_of = {}                     # noise seed changes every time an image is revisited

    def __len__(self):
        return 400 * len(self.de_type)         # dataset_utils.py:142-143

    def _draw_sample(self):
        """The host-side part of TrainDataset.__getitem__: returns (entry, sigma, [(y0, x0, mode)] * 2, de name)."""
        de_num = self.de_type_iterator % len(self.de_type)
        ents = self.entries[de_num]
        if self.de_iterator[de_num] == 0:                                     # :100-104 (element 0 never moves)
            for t in reversed(range(1, len(ents))):
                j = self.rng.randrange(1, t + 1)
                ents[t], ents[j] = ents[j], ents[t]
        e = ents[self.de_iterator[de_num]]
        sigma = 0
        if 'denoising' in self.de_type[de_num]:
            sigma = int(self.de_type[de_num].split('_')[-1])
            if sigma == 0:
                sigma = int(self.np_rng.choice([15, 25, 50]))                # :123-125
        draws = []
        P = self.patch_size
        for _ in range(2):                                                    # :130-131
            y0 = self.rng.randint(0, e['H'] - P)
            x0 = self.rng.randint(0, e['W'] - P)
            draws.append((y0, x0, self.rng.randint(1, 7)))                    # image_utils.py:180
        self.de_iterator[de_num] = (self.de_iterator[de_num] + 1) % len(ents)
        self.de_type_iterator = (self.de_type_iterator + 1) % len(self.de_type)
        return e, sigma, draws, self.de_type[de_num]

    def next_batch(self, B, noise=None):
        """One DataLoader batch: ``([names, de_ids], degrad_patch_1, degrad_patch_2, clean_patch_1, clean_patch_2)``,
        tensors [B,3,P,P] float32 on the device.  ``noise`` ([2B,P,P,3], tests only) overrides the stateless generator."""
        P = self.patch_size
        meta = torch.empty(2 * B, 8, dtype=torch.int64)
        sig = torch.empty(2 * B, dtype=torch.float32)
        names, de_ids = [], []
        for i in range(B):
            e, sigma, draws, de = self._draw_sample()
            visit = self.epoch_of.get(e['uid'], 0)
            self.epoch_of[e['uid']] = visit + 1
            seed = (self.seed * 1000003 + e['uid']) * 1000003 + visit          # one noise field per (image, visit)
            for v, (y0, x0, mode) in enumerate(draws):
                meta[v * B + i] = torch.tensor([e['clean'], e['degraded'], e['H'], e['W'], y0, x0, mode, seed])
                sig[v * B + i] = sigma
            names.append(e['name']); de_ids.append(de)
        deg, clean = ops.crop_augment(self.pool, meta.to(self.device, non_blocking=True),
                                      sig.to(self.device, non_blocking=True), noise, P)
        self.last_meta = meta
        return [names, de_ids], deg[:B], deg[B:], clean[:B], clean[B:]
