"""
Minimal Keras-2.2-compatible front end for the DLWP convolutional nets.  `dlwp_b200.compat.install()` registers it as
`keras` in sys.modules so the reference's example scripts import it unchanged.
"""

from . import backend, callbacks, layers, losses, models, optimizers, regularizers, utils  # noqa: F401
from .engine import Input, Model, Sequential  # noqa: F401

__version__ = '2.2.4-dlwp_b200'
