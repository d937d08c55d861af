"""
dlwp_b200 -- B200-native implementation of the DLWP forecast-rollout hot path (jweyn/DLWP `predict_timeseries`:
periodic-pad + Conv2D stack applied iteratively), behind DLWP's own Python API.

    from dlwp_b200.model import DLWPNeuralNet, DLWPFunctional          # = DLWP.model
    from dlwp_b200.custom import PeriodicPadding2D, RowConnected2D      # = DLWP.custom
    import dlwp_b200.compat; dlwp_b200.compat.install()                 # `import DLWP`, `import keras` aliases

All arithmetic runs in csrc/libdlwp_b200.so (hand-written sm_100a CUDA behind the C ABI of include/dlwp_b200.h).
There is no CPU fallback: importing the compute path without the built library raises.
"""

__version__ = '0.1.0'

from . import _native  # noqa: F401  (does not load the library until first use)
