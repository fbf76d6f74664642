import importlib
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
PKG_NAME = 'frequency-wised_all-in-one_image_restoration_model_b200'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'ref: needs /root/reference (build container only)')


def pytest_collection_modifyitems(config, items):
    has_gpu = torch.cuda.is_available()
    has_ref = os.path.isdir('/root/reference/net')
    for it in items:
        if 'gpu' in it.keywords and not has_gpu:
            it.add_marker(pytest.mark.skip(reason='no CUDA device'))
        if 'ref' in it.keywords and not has_ref:
            it.add_marker(pytest.mark.skip(reason='/root/reference not present'))


@pytest.fixture(scope='session')
def pkg():
    return importlib.import_module(PKG_NAME)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def load_spec(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def t(a):
    return torch.from_numpy(np.asarray(a))
