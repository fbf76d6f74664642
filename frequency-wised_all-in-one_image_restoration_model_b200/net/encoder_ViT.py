from torch import nn


class ViTEncoder(nn.Module):
    def __init__(self, opt):
        raise NotImplementedError('ViTEncoder: pending')
