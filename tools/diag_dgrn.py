import importlib, sys, os, types, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
PKG='frequency-wised_all-in-one_image_restoration_model_b200'
from oracle import detfill
ops = importlib.import_module(PKG + '.ops')
renc_mod = importlib.import_module(PKG + '.net.encoder_ResNet')
dgrn_mod = importlib.import_module(PKG + '.net.decoder_DGRN')
o = types.SimpleNamespace(encoder_type='ResNet', decoder_type='ResNet', encoder_dim=32, L=3, encoder_msa_type='freq', encoder_embed_dim=28, embed_dim=56, degradation_embedding_method=['all_3_bands'], frequency_decompose_type='none', learnable_modulator=False, debug_mode=False, batch_size=2, out_channels=3, batch_wise_decompose=False)
res = {}
for be in (1, 0):
    ops.DEFAULT_GEMM_BACKEND = be
    torch.manual_seed(3)
    renc, dgrn = renc_mod.ResNetEncoder(o), dgrn_mod.DGRN(o)
    detfill.fill_state(renc.state_dict()); detfill.fill_state(dgrn.state_dict())
    x = torch.rand(2, 3, 16, 16); w = torch.randn(2, 3, 16, 16)
    renc, dgrn = renc.cuda().train(), dgrn.cuda().train()
    fk, ok, ik = renc(x.cuda()); yk = dgrn(x.cuda(), ik)
    ((yk * w.cuda()).sum() + ok[0].square().sum()).backward()
    res[be] = ({n: p.grad.clone() for n, p in dgrn.named_parameters()}, yk.detach().clone())
print('output diff', (res[0][1] - res[1][1]).abs().max().item())
bad = 0
for n in res[0][0]:
    a, b = res[0][0][n], res[1][0][n]
    sc = b.abs().max().item()
    err = (a - b).abs()
    if err.max().item() > 2e-3 * sc:
        bad += 1
        print(n, 'max err', err.max().item(), 'scale', sc, 'n>tol', (err > 2e-3 * sc).sum().item(), '/', err.numel(), 'relL2', ((a - b).norm() / b.norm()).item())
print('bad params', bad, 'of', len(res[0][0]))
