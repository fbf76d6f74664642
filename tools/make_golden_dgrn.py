"""Golden vectors for the DGRN-side configurations (BASELINE configs[0] and configs[2]) from the reference's own
modules (build container only; needs /root/reference, ~10 min CPU).

  python tools/make_golden_dgrn.py [cfg0] [vit_train] [resnet_train]

What the reference can and cannot run:
 * ``DCN_layer.forward`` stops at ``assert False`` (deform_conv.py:64; mmcv absent and unpinned): DCNv2 goes through the
   torchvision stand-in of ``tools/ref_shims.patch_dcn`` - **parity unpinned** for that op.
 * ``MoCo.forward`` loops ``range(opt.L)`` over the 1-element ``[out]`` the ResNet / ViT encoders return
   (moco.py:127-128 vs encoder_ResNet.py:47, encoder_ViT.py:203) and dies with IndexError.  The one deviation made
   here - and by the product (net/utils/moco.py) - is ``num_losses = len(out)``: the attribute is set on the reference's
   MoCo object after construction and then its UNMODIFIED forward runs.
 * ViT dropout (p = 0.1, encoder_ViT.py:128-129) is RNG-dependent: the train-step goldens are taken with every
   nn.Dropout at p = 0.

cfg0          ResNetEncoder + DGRN eval forward, sigma = 25, batch 4, 128 x 128 (configs[0] exactly)
vit_train     ViT (4_bands, encoder_dim 64) + DGRN: one train step body (train.py:87-95) on a mixed-degradation batch of 2
resnet_train  ResNet (encoder_dim 256) + DGRN: the same
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import ref_shims  # noqa: E402
from make_golden import spec_of, strided_sample  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
synth = importlib.import_module('frequency-wised_all-in-one_image_restoration_model_b200.synth')
from oracle import detfill  # noqa: E402


def grad_digest_all(net, n_samp=1024):
    out = {}
    for n, p in sorted(net.named_parameters()):
        if p.grad is None:
            continue
        g = p.grad.detach().float().flatten()
        out['gstat/' + n] = np.array([g.sum().item(), g.abs().sum().item(), g.norm().item(), g.abs().max().item()], np.float64)
        out['gsamp/' + n] = strided_sample(g, n_samp).numpy().copy()
    return out


def train_step(opt, tag, kinds):
    from net.model import AirNet
    opt.batch_size = 2
    net = AirNet(opt)
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    json.dump(spec_of(net), open(os.path.join(OUT, f'spec_airnet_{tag}.json'), 'w'))
    detfill.fill_state(net)
    net.train()
    net.E.E.num_losses = 1                       # = len(out) of the ResNet / ViT encoders (module docstring)
    xq, xk, clean = synth.mixed_batch(2, kinds=kinds)
    restored, logits, labels = net(xq, xk)
    ce = sum(torch.nn.functional.cross_entropy(logits[i], labels[i]) for i in range(len(logits))) / len(logits)
    l1 = (restored - clean).abs().mean()
    loss = l1 + 0.6 * ce                                     # train.py:88-92
    loss.backward()
    tr = dict(restored=restored.detach().numpy(), logits=torch.stack(logits).detach().numpy(),
              loss=np.array([loss.item(), l1.item(), ce.item()]), queue=net.E.E.queue.numpy(),
              queue_ptr=net.E.E.queue_ptr.numpy())
    tr.update(grad_digest_all(net))
    sd = net.state_dict()
    for k in sd:
        if k.endswith('running_mean') or k.endswith('running_var'):
            tr['bn/' + k] = sd[k].numpy()
    ks = [k for k in sd if k.startswith('E.E.encoder_k.') and sd[k].is_floating_point() and 'running' not in k]
    for k in (ks[0], ks[len(ks) // 2], ks[-1]):
        tr['kparam/' + k] = strided_sample(sd[k], 4096).numpy().copy()
    np.savez_compressed(os.path.join(OUT, f'airnet_{tag}_train.npz'), **tr)
    print(tag, 'train loss', loss.item(), l1.item(), ce.item(), 'params with grad', sum(1 for k in tr if k.startswith('gstat/')))


def main():
    what = set(sys.argv[1:]) or {'cfg0', 'vit_train', 'resnet_train'}
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_shims.install(['--degradation_embedding_method', 'all_3_bands'])
    from option import options as opt
    ref_shims.patch_dcn()
    opt.decoder_type = 'ResNet'

    if 'cfg0' in what:
        from net.decoder_DGRN import DGRN
        from net.encoder_ResNet import ResNetEncoder
        opt.encoder_type, opt.encoder_dim = 'ResNet', 256
        renc, dgrn = ResNetEncoder(opt), DGRN(opt)
        detfill.fill_state(renc)
        detfill.fill_state(dgrn)
        renc.eval(), dgrn.eval()
        xq, _, clean = synth.noisy_batch(4, 25)
        with torch.no_grad():
            fea, out, inter = renc(xq)
            y = dgrn(xq, inter)
        np.savez_compressed(os.path.join(OUT, 'resnet_dgrn_cfg0.npz'), restored=y.numpy(), fea=fea.numpy(), out=out[0].numpy(),
                            inter_samp=strided_sample(inter, 65536).numpy().copy(),
                            inter_stat=np.array([inter.sum().item(), inter.abs().sum().item()]))
        print('cfg0 restored mean |y|', y.abs().mean().item(), 'PSNR-ish mse', (y - clean).pow(2).mean().item())

    if 'vit_train' in what:
        opt.encoder_type, opt.encoder_dim, opt.frequency_decompose_type = 'ViT', 64, '4_bands'
        train_step(opt, 'vit_dgrn', synth.DEGRADATIONS)
    if 'resnet_train' in what:
        opt.encoder_type, opt.encoder_dim, opt.frequency_decompose_type = 'ResNet', 256, 'none'
        train_step(opt, 'resnet_dgrn', ('sigma25', 'rain'))


if __name__ == '__main__':
    main()
