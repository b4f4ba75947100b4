"""diffusers.utils stand-ins (test-only)."""
import logging as _pylogging
from collections import OrderedDict

import torch

USE_PEFT_BACKEND = False


class BaseOutput(OrderedDict):
    def __init__(self, **kw):
        super().__init__(**kw)
        for k, v in kw.items():
            object.__setattr__(self, k, v)


class logging:  # noqa: N801 - mirrors `from diffusers.utils import logging`
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


def scale_lora_layers(model, scale):
    pass


def unscale_lora_layers(model, scale):
    pass


def deprecate(*args, **kwargs):
    pass


def is_torch_version(op, version):
    from packaging import version as V

    cur = V.parse(torch.__version__.split("+")[0])
    ref = V.parse(version)
    return {">=": cur >= ref, ">": cur > ref, "<": cur < ref, "<=": cur <= ref, "==": cur == ref}[op]
