import logging as _logging
import operator
from collections import OrderedDict

import torch
from packaging import version

_OPS = {">": operator.gt, ">=": operator.ge, "==": operator.eq, "!=": operator.ne, "<=": operator.le, "<": operator.lt}
USE_PEFT_BACKEND = False
WEIGHTS_NAME = "diffusion_pytorch_model.bin"


def is_torch_version(op: str, ver: str) -> bool:
    return _OPS[op](version.parse(version.parse(torch.__version__).base_version), version.parse(ver))


def is_accelerate_available() -> bool:
    return False


class BaseOutput(OrderedDict):
    pass


class _Logging:
    @staticmethod
    def get_logger(name):
        return _logging.getLogger(name)


logging = _Logging()


# ---- names the reference PIPELINE files import (pipeline_easyanimate.py:33-36)
BACKENDS_MAPPING = {}


def deprecate(*args, **kwargs):
    pass


def is_bs4_available() -> bool:
    return False


def is_ftfy_available() -> bool:
    return False


def is_torch_xla_available() -> bool:
    return False


def replace_example_docstring(example_docstring):
    def wrap(fn):
        return fn

    return wrap
