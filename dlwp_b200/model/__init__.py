"""Mirror of `DLWP.model` (reference DLWP/model/__init__.py:11-17).  The xarray/netCDF-dependent classes are resolved
lazily because those packages are unavailable offline and lie outside the rollout hot path."""

from .generators import (ArrayDataGenerator, ArraySeriesGenerator, DataGenerator, DeviceSeriesGenerator,  # noqa: F401
                         SeriesDataGenerator, SmartDataGenerator)
from .extensions import TimeSeriesEstimator  # noqa: F401,E402
from .models import DLWPFunctional, DLWPNeuralNet  # noqa: F401,E402
from .models_torch import DLWPTorchNN  # noqa: F401,E402

_OUT_OF_SCOPE = {
    'Preprocessor': 'DLWP/model/preprocessing.py (offline data preparation, needs netCDF4/xarray)',
    'verify': 'DLWP/model/verify.py (post-processing metrics, needs xarray/pandas)',
}


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError('DLWP.model.%s is outside the rollout hot path that dlwp_b200 implements: %s' %
                                  (name, _OUT_OF_SCOPE[name]))
    raise AttributeError(name)
