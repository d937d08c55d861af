"""keras.backend names touched by the example scripts (examples/validate.py:11,245)."""

from .engine import clear_session  # noqa: F401
from .layers import _norm_data_format as normalize_data_format  # noqa: F401


def backend():
    return 'dlwp_b200'


def image_data_format():
    return 'channels_last'


def floatx():
    return 'float32'
