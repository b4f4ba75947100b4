from dataclasses import dataclass

from torch import nn


def zero_module(module):
    for p in module.parameters():
        nn.init.zeros_(p)
    return module


class ControlNetConditioningEmbedding(nn.Module):
    pass


@dataclass
class ControlNetOutput:
    down_block_res_samples: tuple = None
    mid_block_res_sample: object = None
