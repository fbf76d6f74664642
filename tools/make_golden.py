"""Generate ``tests/golden/*`` from the UNMODIFIED reference (build container only).

Run:  python tools/make_golden.py            (needs /root/reference; ~2 min CPU)

The reference holds no golden vectors of its own, so the oracle is pinned to
outputs of the reference's modules executed here: parameters come from
``oracle.detfill.fill_state`` (pure function of name+shape), inputs from the package's ``synth``
module (pure function of a seed), and what is committed is the reference's *outputs*, small
enough for git.  ``tests/test_oracle_golden.py`` replays them through ``oracle/`` (CPU) and
``tests/test_gpu_golden.py`` through the CUDA path.
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

OUT = os.path.join(ROOT, 'tests', 'golden')
synth = importlib.import_module('frequency-wised_all-in-one_image_restoration_model_b200.synth')
from oracle import detfill  # noqa: E402


def spec_of(module):
    return {k: [list(v.shape), str(v.dtype).replace('torch.', '')] for k, v in module.state_dict().items()}


def strided_sample(t, n):
    """Up to n elements of t.flatten() at a fixed stride (a pure function of numel and n: tests recompute the indices)."""
    f = t.flatten()
    step = max(1, -(-f.numel() // n))
    return f[::step]


def grad_digest(named_params, every=53, n_all=256, n_big=4096):
    """For EVERY parameter with a gradient: gstat = [sum, abs-sum, l2 norm, max-abs] and gsamp = a strided sample of up to
    ``n_all`` elements (``n_big`` for every ``every``-th parameter, the relative-position tables and the lambda MLPs), so
    the GPU test can bound the max-abs gradient error per parameter over the whole model.  grad_head / grad_sum are the
    round-1 keys (first 256 entries of every ``every``-th parameter), kept for the CPU oracle test."""
    out = {}
    byname = dict(named_params)
    names = sorted(n for n, p in named_params if p.grad is not None)
    for j, n in enumerate(names):
        g = byname[n].grad.detach().float().flatten()
        big = not (j % every and 'relative_position_bias_table' not in n and '.mlp.1.0.' not in n)
        out['gstat/' + n] = np.array([g.sum().item(), g.abs().sum().item(), g.norm().item(), g.abs().max().item()], np.float64)
        out['gsamp/' + n] = strided_sample(g, n_big if big else n_all).numpy().copy()
        if big:
            out['grad_sum/' + n] = np.array([g.sum().item(), g.abs().sum().item()], np.float64)
            out['grad_head/' + n] = g[:256].numpy().copy()
    return out


def name_droppaths(net):
    for name, m in net.named_modules():
        if type(m).__name__ == 'DropPath':
            m._fa_name = name
    import timm.models.layers as L
    orig = L.DropPath.forward

    def fwd(self, x):
        n0 = len(ref_shims.SCALES_LOG)
        y = orig(self, x)
        if len(ref_shims.SCALES_LOG) > n0:
            ref_shims.SCALES_LOG[-1] = (self._fa_name, ref_shims.SCALES_LOG[-1])
        return y
    L.DropPath.forward = fwd


def dp_to_npz(log):
    seen, out = {}, {}
    for name, r in log:
        blk = name[:-len('.drop_path')]
        i = seen.get(blk, 0)
        seen[blk] = i + 1
        out[f'dp/{blk}/{i}'] = r.numpy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ------------------------------------------------------------------ A. frequency decomposition
    ref_shims.install(['--degradation_embedding_method', 'all_3_bands'])
    from net.utils.frequency_decompose import FrequencyDecompose
    from option import options as opt
    fx = {}
    g = torch.Generator().manual_seed(7)
    x64 = torch.rand(2, 3, 64, 64, generator=g).softmax(-1)
    x128 = torch.rand(1, 3, 128, 128, generator=g)
    x16 = torch.rand(2, 2, 16, 16, generator=g)
    fx['x64'], fx['x128'], fx['x16'] = x64.numpy(), x128.numpy(), x16.numpy()
    cases = [('frequency_decompose_1', 0.5, 'x64', True), ('frequency_decompose_1', 1.0, 'x64', True),
             ('frequency_decompose_1', 0.5, 'x128', True), ('frequency_decompose', 0.25, 'x64', True),
             ('frequency_decompose', 0.125, 'x64', True), ('frequency_decompose', 0.5, 'x64', True),
             ('frequency_decompose', 0.25, 'x128', False), ('frequency_decompose', 1.0, 'x16', 'visual'),
             ('frequency_decompose_dc', 0.5, 'x64', True), ('frequency_decompose_1', 0.5, 'x16', True)]
    meta = []
    for n, (kind, size, xn, inv) in enumerate(cases):
        x = torch.from_numpy(fx[xn])
        y = FrequencyDecompose(kind, size, x.shape[-2], x.shape[-1], inverse=inv)(x)
        fx[f'out{n}'] = y.numpy()
        meta.append([kind, size, xn, inv if isinstance(inv, str) else bool(inv)])
    fx['meta'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, 'freq.npz'), **fx)
    print('freq.npz', len(cases), 'cases')

    # ------------------------------------------------------------------ B. AirNet Uformer+Uformer
    from net.model import AirNet
    opt.batch_size = 2
    net = AirNet(opt)
    json.dump(spec_of(net), open(os.path.join(OUT, 'spec_airnet_uformer_uformer_L3.json'), 'w'))
    detfill.fill_state(net)
    name_droppaths(net)
    xq, xk, clean = synth.noisy_batch(2, 25)

    net.eval()
    with torch.no_grad():
        restored = net(xq[:1], xq[:1])
        fea, inter = net.E(xq[:1], xq[:1])
    np.savez_compressed(os.path.join(OUT, 'airnet_uu_eval.npz'), restored=restored.numpy(),
                        inter=torch.stack(inter).numpy())
    print('eval restored', restored.abs().mean().item())

    net.train()
    ref_shims.SCALES_LOG.clear()
    ref_shims.SCALES_RNG.manual_seed(99)
    restored, logits, labels = net(xq, xk)
    ce = sum(torch.nn.functional.cross_entropy(logits[i], labels[i]) for i in range(3)) / 3
    l1 = (restored - clean).abs().mean()
    loss = l1 + 0.6 * ce                                     # train.py:88-92 with the L=3 weight of option.py:59-60
    loss.backward()
    tr = dict(restored=restored.detach().numpy(), logits=torch.stack(logits).detach().numpy(),
              loss=np.array([loss.item(), l1.item(), ce.item()]),
              queue=net.E.E.queue.numpy(), queue_ptr=net.E.E.queue_ptr.numpy())
    tr.update(dp_to_npz(ref_shims.SCALES_LOG))
    tr.update(grad_digest(list(net.named_parameters())))
    sd = net.state_dict()
    for k in sd:
        if k.endswith('running_mean') or k.endswith('running_var'):
            tr['bn/' + k] = sd[k].numpy()
    # a few momentum-updated key-encoder tensors
    for k in ['E.E.encoder_k.uformer.input_proj.proj.0.weight', 'E.E.encoder_k.mlp.2.2.weight']:
        tr['kparam/' + k] = sd[k].numpy()
    np.savez_compressed(os.path.join(OUT, 'airnet_uu_train.npz'), **tr)
    print('train loss', loss.item(), l1.item(), ce.item(), 'dp draws', len(ref_shims.SCALES_LOG))
    del net

    # ------------------------------------------------------------------ C. other variants, module level
    for method, L in (('all_DC', 3), ('all_2_bands', 2)):
        opt.degradation_embedding_method = [method]
        opt.L = L
        from net.decoder_Uformer import UformerDecoder
        dec = UformerDecoder(opt)
        detfill.fill_state(dec)
        dec.eval()
        gi = torch.Generator().manual_seed(11)
        inter = tuple(torch.randn(1, 64, 448, generator=gi) for _ in range(L))
        with torch.no_grad():
            y = dec(xq[:1], inter)
        np.savez_compressed(os.path.join(OUT, f'dec_{method}.npz'), restored=y.numpy(),
                            inter=torch.stack(inter).numpy())
        json.dump(spec_of(dec), open(os.path.join(OUT, f'spec_dec_{method}.json'), 'w'))
        print('dec', method, y.abs().mean().item())
        del dec
    opt.L = 3
    opt.degradation_embedding_method = ['all_3_bands']
    opt.encoder_msa_type = 'origin'
    from net.encoder_Uformer import UformerEncoder
    enc = UformerEncoder(opt)
    detfill.fill_state(enc)
    enc.eval()
    with torch.no_grad():
        _, out, inter = enc(xq[:1])
    np.savez_compressed(os.path.join(OUT, 'enc_origin_eval.npz'), out=torch.stack(out).numpy(),
                        inter=torch.stack(inter).numpy())
    json.dump(spec_of(enc), open(os.path.join(OUT, 'spec_enc_origin.json'), 'w'))
    opt.encoder_msa_type = 'freq'
    del enc

    # ResNet encoder + DGRN (BASELINE config 1): DCN through the torchvision stand-in
    ref_shims.patch_dcn()
    from net.encoder_ResNet import ResNetEncoder
    from net.decoder_DGRN import DGRN
    opt.encoder_type, opt.encoder_dim = 'ResNet', 256
    renc, dgrn = ResNetEncoder(opt), DGRN(opt)
    json.dump(spec_of(renc), open(os.path.join(OUT, 'spec_resnet_encoder.json'), 'w'))
    json.dump(spec_of(dgrn), open(os.path.join(OUT, 'spec_dgrn64.json'), 'w'))
    detfill.fill_state(renc)
    detfill.fill_state(dgrn)
    renc.eval(), dgrn.eval()
    x1 = xq[:1, :, :64, :64].contiguous()
    with torch.no_grad():
        fea, out, inter = renc(x1)
        y = dgrn(x1, inter)
    res = dict(fea=fea.numpy(), out=out[0].numpy(), inter=inter.numpy(), restored=y.numpy())
    renc.train()
    fea, out, inter = renc(xq)
    res.update(train_fea=fea.detach().numpy(), train_out=out[0].detach().numpy())
    np.savez_compressed(os.path.join(OUT, 'resnet_dgrn.npz'), **res)
    print('resnet+dgrn', y.abs().mean().item())

    # ViT encoder, 4 bands, encoder_dim 64 (the authors' own ViT runs, plot_LFS_distribution.py:26)
    from net.encoder_ViT import ViTEncoder
    opt.encoder_type, opt.encoder_dim, opt.frequency_decompose_type = 'ViT', 64, '4_bands'
    vit = ViTEncoder(opt)
    json.dump(spec_of(vit), open(os.path.join(OUT, 'spec_vit_encoder_ed64.json'), 'w'))
    detfill.fill_state(vit)
    vit.eval()
    with torch.no_grad():
        fea, out, inter = vit(xq.clone())
    np.savez_compressed(os.path.join(OUT, 'vit_encoder.npz'), fea=fea.numpy(), out=out[0].numpy(),
                        inter_head=inter[:, :4, :8, :].numpy(), inter_sum=np.array([inter.sum().item(), inter.abs().sum().item()]))
    print('vit', fea.abs().mean().item())


if __name__ == '__main__':
    main()
