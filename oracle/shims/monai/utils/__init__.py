"""monai.utils names imported at inference/sliding_window_inferer.py:20."""
from enum import Enum


class BlendMode(str, Enum):
    CONSTANT = "constant"
    GAUSSIAN = "gaussian"


class PytorchPadMode(str, Enum):
    CONSTANT = "constant"
    REFLECT = "reflect"
    REPLICATE = "replicate"
    CIRCULAR = "circular"


def fall_back_tuple(user_provided, default, func=lambda x: x and x > 0):
    nd = len(default)
    if isinstance(user_provided, int):
        user_provided = (user_provided,) * nd
    return tuple(u if func(u) else d for u, d in zip(user_provided, default))


def look_up_option(opt, supported, default="no_default"):
    if isinstance(supported, type) and issubclass(supported, Enum):
        return supported(opt)
    if opt in supported:
        return opt
    if default != "no_default":
        return default
    raise ValueError(opt)
