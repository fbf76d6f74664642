"""GPU: the CUDA modules reproduce the committed outputs of the real reference (tests/golden/*) and the
oracle's gradients, with the same name-keyed weights and seeded inputs.  Tolerance: north_star's 1e-3 max-abs
on activations and gradients (fp32 kernels sit far inside it)."""
import importlib
import types

import pytest
import torch
import torch.nn.functional as F

from conftest import PKG_NAME, load_golden, load_spec, t
from oracle import detfill

pytestmark = pytest.mark.gpu
synth = importlib.import_module(PKG_NAME + '.synth')


def make_opt(**kw):
    o = types.SimpleNamespace(encoder_type='Uformer', decoder_type='Uformer', encoder_dim=256, L=3,
                              encoder_msa_type='freq', encoder_embed_dim=28, embed_dim=56,
                              degradation_embedding_method=['all_3_bands'], frequency_decompose_type='none',
                              learnable_modulator=False, debug_mode=False, batch_size=2, out_channels=3,
                              batch_wise_decompose=False)
    o.__dict__.update(kw)
    return o


def load_det(module, spec_name):
    spec = load_spec(spec_name)
    sd = module.state_dict()
    assert set(sd.keys()) == set(spec.keys()), (sorted(set(spec) - set(sd))[:5], sorted(set(sd) - set(spec))[:5])
    for k, v in sd.items():
        assert list(v.shape) == spec[k][0], k
    detfill.fill_state(sd)        # in place on the module's own tensors (CPU), then move
    return module


def maxerr(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


@pytest.mark.parametrize('method,L', [('all_DC', 3), ('all_2_bands', 2)])
def test_decoder_variants(method, L):
    dec_mod = importlib.import_module(PKG_NAME + '.net.decoder_Uformer')
    g = load_golden(f'dec_{method}.npz')
    dec = load_det(dec_mod.UformerDecoder(make_opt(degradation_embedding_method=[method], L=L)), f'spec_dec_{method}.json')
    dec = dec.cuda().eval()
    xq, _, _ = synth.noisy_batch(2, 25)
    with torch.no_grad():
        y = dec(xq[:1].cuda(), tuple(t(g['inter']).cuda()))
    assert maxerr(y, t(g['restored'])) < 1e-3


def test_encoder_origin_msa():
    enc_mod = importlib.import_module(PKG_NAME + '.net.encoder_Uformer')
    g = load_golden('enc_origin_eval.npz')
    enc = load_det(enc_mod.UformerEncoder(make_opt(encoder_msa_type='origin')), 'spec_enc_origin.json').cuda().eval()
    xq, _, _ = synth.noisy_batch(2, 25)
    with torch.no_grad():
        _, out, inter = enc(xq[:1].cuda())
    assert maxerr(torch.stack(inter), t(g['inter'])) < 1e-3
    assert maxerr(torch.stack(out), t(g['out'])) < 1e-3


@pytest.fixture(scope='module')
def airnet():
    model = importlib.import_module(PKG_NAME + '.net.model')
    net = load_det(model.AirNet(make_opt()), 'spec_airnet_uformer_uformer_L3.json')
    return net.cuda()


def test_airnet_eval(airnet):
    g = load_golden('airnet_uu_eval.npz')
    airnet.eval()
    xq, _, _ = synth.noisy_batch(2, 25)
    with torch.no_grad():
        y = airnet(xq[:1].cuda(), xq[:1].cuda())
        _, _, inter = airnet.E.E.encoder_q(xq[:1].cuda())
    assert maxerr(torch.stack(inter), t(g['inter'])) < 1e-3
    assert maxerr(y, t(g['restored'])) < 1e-3


def test_airnet_train_step(airnet):
    losses = importlib.import_module(PKG_NAME + '.losses')
    g = load_golden('airnet_uu_train.npz')
    # the fixture may have been used in eval before: restore the name-keyed state
    detfill.fill_state(airnet.state_dict())
    airnet.train()
    dp = {}
    for k in g:
        if k.startswith('dp/'):
            _, blk, i = k.split('/')
            dp.setdefault(blk, [None, None])[int(i)] = t(g[k]).cuda()
    for blk, (s0, s1) in dp.items():
        airnet.get_submodule(blk).forced_dp = (s0, s1)
    xq, xk, clean = (v.cuda() for v in synth.noisy_batch(2, 25))
    restored, logits, labels = airnet(xq, xk)
    assert maxerr(restored, t(g['restored'])) < 1e-3
    assert maxerr(torch.stack(logits), t(g['logits'])) < 5e-3
    ce = sum(F.cross_entropy(logits[i], labels[i]) for i in range(3)) / 3
    l1 = losses.l1_loss(restored, clean)
    loss = l1 + 0.6 * ce
    assert abs(loss.item() - g['loss'][0]) < 1e-3
    airnet.zero_grad()
    loss.backward()
    params = dict(airnet.named_parameters())
    checked, worst = 0, 0.0
    for k in g:
        if k.startswith('grad_head/'):
            name = k[len('grad_head/'):]
            gr = params[name].grad.flatten()[:256].cpu()
            ref = t(g[k])
            scale = max(ref.abs().max().item(), 1e-6)
            err = (gr - ref).abs().max().item()
            worst = max(worst, err / scale)
            assert err <= 1e-3 * max(scale, 1.0) or err <= 5e-3 * scale, (name, err, scale)
            s = g['grad_sum/' + name]
            assert abs(params[name].grad.abs().sum().item() - s[1]) <= 5e-3 * s[1] + 1e-6, name
            checked += 1
    assert checked > 20
    sd = airnet.state_dict()
    for k in g:
        if k.startswith('kparam/'):
            assert maxerr(sd[k[len('kparam/'):]], t(g[k])) < 1e-6, k
        if k.startswith('bn/'):
            assert maxerr(sd[k[3:]], t(g[k])) < 1e-4, k
    assert maxerr(sd['E.E.queue'], t(g['queue'])) < 1e-4
    assert int(sd['E.E.queue_ptr']) == int(g['queue_ptr'][0])
    for blk in dp:
        airnet.get_submodule(blk).forced_dp = None
