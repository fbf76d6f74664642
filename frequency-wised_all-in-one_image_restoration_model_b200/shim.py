"""``install()`` makes ``import net...`` resolve to this package's module tree, so the reference's own ``train.py`` /
``test.py`` / plot scripts (``from net.model import AirNet``, train.py:15; ``from net.utils.frequency_decompose import
FrequencyDecompose``, train.py:16) run unchanged on the CUDA-backed classes.

Put ``import importlib; importlib.import_module("frequency-wised_all-in-one_image_restoration_model_b200.shim").install()``
in ``sitecustomize.py`` or at the top of the script, before the first ``import net``."""
import importlib
import sys

SUBMODULES = ('model', 'encoder_Uformer', 'decoder_Uformer', 'encoder_ResNet', 'encoder_ViT', 'decoder_DGRN', 'utils',
              'utils.frequency_decompose', 'utils.moco', 'utils.leff', 'utils.deform_conv')


def install():
    pkg = __name__.rsplit('.', 1)[0]
    if 'net' in sys.modules and not sys.modules['net'].__name__.startswith(pkg):
        raise RuntimeError('freqair shim: a different `net` package is already imported; install() must run first')
    sys.modules['net'] = importlib.import_module(pkg + '.net')
    for sub in SUBMODULES:
        sys.modules['net.' + sub] = importlib.import_module(pkg + '.net.' + sub)
    return sys.modules['net']
