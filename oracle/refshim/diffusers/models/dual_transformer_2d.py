from torch import nn


class DualTransformer2DModel(nn.Module):
    pass
