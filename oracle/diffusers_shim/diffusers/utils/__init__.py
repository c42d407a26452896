import logging as _pylogging
from packaging import version as _version
import torch as _torch

USE_PEFT_BACKEND = False
WEIGHTS_NAME = "diffusion_pytorch_model.bin"
SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"


class _Logging:
    @staticmethod
    def get_logger(name=None):
        return _pylogging.getLogger(name)


logging = _Logging()


def deprecate(*args, **kwargs):
    return None


def is_torch_version(op, ver):
    cur = _version.parse(_torch.__version__.split("+")[0])
    ref = _version.parse(ver)
    return {">=": cur >= ref, ">": cur > ref, "<=": cur <= ref, "<": cur < ref, "==": cur == ref}[op]


def scale_lora_layers(model, weight):
    return None


def unscale_lora_layers(model, weight=None):
    return None
