from torch import nn


class AdaGroupNorm(nn.Module):
    pass
