"""CPU: the drop-in boundary itself - the C-ABI library loads and exports every symbol include/freqair.h declares (no
compute is launched), the host-side modules carry the reference's state_dict keys and shapes, ops refuse CPU tensors
(there is no CPU path), and configurations that crash at reference HEAD raise at construction."""
import ctypes
import importlib
import os
import types

import pytest
import torch

from conftest import PKG_NAME, ROOT, load_spec


def make_opt(**kw):
    o = types.SimpleNamespace(encoder_type='Uformer', decoder_type='Uformer', encoder_dim=256, L=3,
                              encoder_msa_type='freq', encoder_embed_dim=28, embed_dim=56,
                              degradation_embedding_method=['all_3_bands'], frequency_decompose_type='none',
                              learnable_modulator=False, debug_mode=False, batch_size=2, out_channels=3,
                              batch_wise_decompose=False)
    o.__dict__.update(kw)
    return o


def test_library_exports_every_declared_symbol():
    lib_mod = importlib.import_module(PKG_NAME + '._lib')
    if not os.path.exists(lib_mod.LIB_PATH):
        importlib.import_module(PKG_NAME + '.build').build()          # nvcc cross-compiles without a GPU
    protos = lib_mod.parse_header()
    assert len(protos) >= 50
    raw = ctypes.CDLL(lib_mod.LIB_PATH)
    for name in protos:
        assert hasattr(raw, name), f'{name} is declared in include/freqair.h but not exported'
    lib = lib_mod.load()
    assert lib.fa_version().decode().startswith('freqair')
    assert lib.fa_last_error_string() is not None
    # the header is the single source of truth: every exported fa_* symbol is declared there
    import subprocess
    out = subprocess.run(['nm', '-D', '--defined-only', lib_mod.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if ' T fa_' in l}
    internal = {'fa_set_error', 'fa_count_launch', 'fa_gemm_simt_launch', 'fa_gemm_tc_launch', 'fa_a_rowsum', 'fa_a_rowsum_impl'}
    assert {e for e in exported if not e.startswith('_Z')} - internal <= set(protos), sorted(exported - set(protos) - internal)[:5]


@pytest.mark.parametrize('build,spec', [
    (lambda m: m('net.model').AirNet(make_opt()), 'spec_airnet_uformer_uformer_L3.json'),
    (lambda m: m('net.decoder_Uformer').UformerDecoder(make_opt(degradation_embedding_method=['all_DC'])), 'spec_dec_all_DC.json'),
    (lambda m: m('net.encoder_Uformer').UformerEncoder(make_opt(encoder_msa_type='origin')), 'spec_enc_origin.json'),
    (lambda m: m('net.encoder_ResNet').ResNetEncoder(make_opt(encoder_type='ResNet', decoder_type='ResNet')), 'spec_resnet_encoder.json'),
    (lambda m: m('net.decoder_DGRN').DGRN(make_opt(encoder_type='ResNet', decoder_type='ResNet')), 'spec_dgrn64.json'),
    (lambda m: m('net.encoder_ViT').ViTEncoder(make_opt(encoder_type='ViT', encoder_dim=64, frequency_decompose_type='4_bands')), 'spec_vit_encoder_ed64.json'),
])
def test_state_dict_keys_and_shapes_match_the_reference(build, spec):
    """tests/golden/spec_*.json were dumped from the reference's own modules (tools/make_golden.py)."""
    mod = build(lambda name: importlib.import_module(PKG_NAME + '.' + name))
    sd = mod.state_dict()
    ref = load_spec(spec)
    assert set(sd) == set(ref), (sorted(set(ref) - set(sd))[:3], sorted(set(sd) - set(ref))[:3])
    for k, v in sd.items():
        assert list(v.shape) == ref[k][0], k


def test_ops_refuse_cpu_tensors():
    ops = importlib.import_module(PKG_NAME + '.ops')
    x = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match='no CPU path|CUDA'):
        ops.layernorm_fwd(x, torch.ones(8), torch.zeros(8))
    with pytest.raises(RuntimeError):
        ops.gemm(x, torch.randn(3, 8), torch.empty(4, 3))


@pytest.mark.parametrize('kw', [dict(degradation_embedding_method=['residual']), dict(debug_mode=True),
                                dict(frequency_decompose_type='2_bands'), dict(learnable_modulator=True)])
def test_configurations_that_crash_at_reference_head_raise(kw):
    dec = importlib.import_module(PKG_NAME + '.net.decoder_Uformer')
    with pytest.raises(NotImplementedError):
        dec.UformerDecoder(make_opt(**kw))


def test_dgrn_with_uformer_encoder_uses_the_documented_adapter():
    """BASELINE configs[3] pairs the Uformer encoder with DGRN; the reference crashes on it (decoder_DGRN.py:120-129).
    The package defines the pairing: n_feats = encoder_dim // 4, state_dict of the ResNet-encoder DGRN, degradation map
    = band mean -> first n_feats channels -> nearest upsampling of the 8 x 8 token grid (uformer_inter_to_map)."""
    dgrn = importlib.import_module(PKG_NAME + '.net.decoder_DGRN')
    d = dgrn.DGRN(make_opt(encoder_type='Uformer', decoder_type='ResNet'))
    ref = load_spec('spec_dgrn64.json')
    assert {k: list(v.shape) for k, v in d.state_dict().items()} == {k: v[0] for k, v in ref.items()}
    inter = tuple(torch.arange(2 * 64 * 448, dtype=torch.float32).view(2, 64, 448) * (i + 1) for i in range(3))
    m = dgrn.uformer_inter_to_map(inter, 64, 128, 128)
    assert m.shape == (2, 64, 128, 128)
    mean = (inter[0] + inter[1] + inter[2]) / 3
    assert torch.equal(m[1, 5, 17, 100], mean[1, (17 // 16) * 8 + 100 // 16, 5])
    assert torch.equal(m[:, :, 0:16, 16:32], m[:, :, 0:1, 16:17].expand(-1, -1, 16, 16))


def test_tile_origins_follow_test_py():
    synth = importlib.import_module(PKG_NAME + '.synth')
    assert synth.tile_indices(512, 512) == ([0, 128, 256, 384], [0, 128, 256, 384])
    assert synth.tile_indices(200, 300) == ([0, 72], [0, 128, 172])
    assert synth.tile_indices(128, 128) == ([0], [0])
