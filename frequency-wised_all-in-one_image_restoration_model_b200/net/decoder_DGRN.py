from torch import nn


class DGRN(nn.Module):
    def __init__(self, opt):
        raise NotImplementedError('DGRN: pending')
