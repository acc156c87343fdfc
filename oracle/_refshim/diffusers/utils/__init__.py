import operator

import torch
from packaging import version

_OPS = {">": operator.gt, ">=": operator.ge, "==": operator.eq, "!=": operator.ne, "<=": operator.le, "<": operator.lt}


def is_torch_version(op: str, ver: str) -> bool:
    return _OPS[op](version.parse(version.parse(torch.__version__).base_version), version.parse(ver))
