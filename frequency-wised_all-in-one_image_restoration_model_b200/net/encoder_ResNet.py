from torch import nn


class ResNetEncoder(nn.Module):
    def __init__(self, opt):
        raise NotImplementedError('ResNetEncoder: pending')
