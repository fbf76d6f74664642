"""CPU: INTEGRATION.md section 3 - with the ``sys.modules['net']`` shim installed, the imports the reference's scripts
perform resolve to this package, ``AirNet(opt)`` builds from an option namespace, and ``optim.Adam(net.parameters())``
(train.py:63) sees the parameter list (names, order, shapes) the reference's own AirNet exposes.  The second test runs
the reference's unmodified ``option.py`` and ``net/model.py`` side by side (build container only)."""
import json
import os
import subprocess
import sys
import types

import pytest

from conftest import PKG_NAME, ROOT, load_spec

OURS = r'''
import importlib, json, sys, types
sys.path.insert(0, {root!r})
importlib.import_module({pkg!r} + '.shim').install()
{get_opt}
from net.model import AirNet                                   # train.py:15 / test.py:13
from net.utils.frequency_decompose import FrequencyDecompose   # train.py:16
import torch
net = AirNet(opt)
opt_ = torch.optim.Adam(net.parameters(), lr=opt.lr if getattr(opt, 'lr', None) else 2e-4)   # train.py:63
n_opt = sum(len(g['params']) for g in opt_.param_groups)
named = [(k, list(v.shape)) for k, v in net.named_parameters()]
assert n_opt == len(named)
print(json.dumps(dict(module=AirNet.__module__, fd=FrequencyDecompose.__module__, params=named,
                      state=[(k, list(v.shape)) for k, v in net.state_dict().items()])))
'''

REF = r'''
import json, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + '/tools')
import ref_shims
ref_shims.install(['--degradation_embedding_method', 'all_3_bands'])
from option import options as opt
from net.model import AirNet
net = AirNet(opt)
print(json.dumps(dict(params=[(k, list(v.shape)) for k, v in net.named_parameters()],
                      state=[(k, list(v.shape)) for k, v in net.state_dict().items()], batch=opt.batch_size)))
'''

NS_OPT = '''opt = types.SimpleNamespace(encoder_type='Uformer', decoder_type='Uformer', encoder_dim=256, L=3, encoder_msa_type='freq',
    encoder_embed_dim=28, embed_dim=56, degradation_embedding_method=['all_3_bands'], frequency_decompose_type='none',
    learnable_modulator=False, debug_mode=False, batch_size=2, out_channels=3, batch_wise_decompose=False)'''
REF_OPT = '''sys.path.insert(0, '/root/reference'); sys.argv = ['train.py', '--degradation_embedding_method', 'all_3_bands']
from option import options as opt                              # the reference's own, unmodified option.py'''


def _run(code):
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_shim_resolves_net_to_this_package():
    out = _run(OURS.format(root=ROOT, pkg=PKG_NAME, get_opt=NS_OPT))
    assert out['module'].startswith(PKG_NAME) and out['fd'].startswith(PKG_NAME)
    spec = load_spec('spec_airnet_uformer_uformer_L3.json')          # dumped from the reference's AirNet
    assert [k for k, _ in out['state']] == list(spec.keys())           # same names in the same order
    assert all(shape == spec[k][0] for k, shape in out['state'])


@pytest.mark.ref
def test_shim_with_reference_option_py_matches_reference_parameter_list():
    ours = _run(OURS.format(root=ROOT, pkg=PKG_NAME, get_opt=REF_OPT))
    ref = _run(REF.format(root=ROOT))
    assert ours['params'] == ref['params']                 # what optim.Adam(net.parameters()) iterates over
    assert ours['state'] == ref['state']
