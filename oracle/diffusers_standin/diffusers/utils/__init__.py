import logging as _logging
from collections import OrderedDict
from dataclasses import fields, is_dataclass

WEIGHTS_NAME = "diffusion_pytorch_model.bin"
SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"


class BaseOutput(OrderedDict):
    """dataclass-style output that also indexes like a tuple (diffusers.utils.BaseOutput)."""

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)

    def __post_init__(self):
        if is_dataclass(self):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    OrderedDict.__setitem__(self, f.name, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return OrderedDict.__getitem__(self, k)
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(OrderedDict.__getitem__(self, k) for k in self.keys())


class _Logging:
    @staticmethod
    def get_logger(name):
        return _logging.getLogger(name)


logging = _Logging()

USE_PEFT_BACKEND = False


def deprecate(*args, **kwargs):
    return None


def scale_lora_layers(model, weight):
    return None


def unscale_lora_layers(model, weight=None):
    return None


def is_torch_version(op, version):
    import operator

    import torch
    from packaging.version import parse
    ops = {">": operator.gt, ">=": operator.ge, "==": operator.eq, "<": operator.lt, "<=": operator.le}
    return ops[op](parse(parse(torch.__version__).base_version), parse(version))
