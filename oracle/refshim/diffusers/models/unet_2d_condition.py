from dataclasses import dataclass

import torch


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor = None
