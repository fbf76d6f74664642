"""AirNet / Encoder / Decoder with the reference's assembly contract (net/model.py:13-71): encoder and decoder
classes are resolved by name from ``opt.encoder_type`` / ``opt.decoder_type``; train mode returns
``(restored, logits, labels)``, eval mode ``restored``."""
from torch import nn

from .decoder_DGRN import DGRN as ResNetDecoder            # noqa: F401  (looked up through globals())
from .decoder_Uformer import UformerDecoder                # noqa: F401
from .encoder_ResNet import ResNetEncoder                  # noqa: F401
from .encoder_Uformer import UformerEncoder                # noqa: F401
from .encoder_ViT import ViTEncoder                        # noqa: F401
from .utils.moco import MoCo


class Decoder(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.R = globals()[opt.decoder_type + 'Decoder'](opt)

    def forward(self, x_query, inter):
        return self.R(x_query, inter)


class Encoder(nn.Module):
    def __init__(self, opt):
        super().__init__()
        encoder = globals()[opt.encoder_type + 'Encoder']
        self.E = MoCo(opt=opt, base_encoder=encoder, dim=opt.encoder_dim, K=opt.batch_size * 3)

    def forward(self, x_query, x_key):
        if self.training:
            return self.E(x_query, x_key)
        return self.E(x_query, x_query)


class AirNet(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.R = Decoder(opt)
        self.E = Encoder(opt)

    def forward(self, x_query, x_key):
        if self.training:
            fea, logits, labels, inter = self.E(x_query, x_key)
            return self.R(x_query, inter), logits, labels
        fea, inter = self.E(x_query, x_query)
        return self.R(x_query, inter)
