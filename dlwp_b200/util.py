"""
Mirror of the parts of `DLWP.util` on or next to the hot path (reference DLWP/util.py): the name -> class registry used
by `build_model`, and model save / load.
"""

import pickle
from copy import copy
from importlib import import_module

import numpy as np

# module names the reference's scripts and build_model use -> where they live in this package
_MODULE_ALIASES = {
    'keras.layers': 'dlwp_b200.keras.layers',
    'keras.models': 'dlwp_b200.keras.models',
    'keras.callbacks': 'dlwp_b200.keras.callbacks',
    'DLWP.custom': 'dlwp_b200.custom',
    'DLWP.util': 'dlwp_b200.util',
}


def get_from_class(module_name, class_name):
    """
    `from module_name import class_name` by strings: DLWP/util.py:82-93.  This is the layer-plugin registry of
    DLWPNeuralNet.build_model (models.py:97-103); `keras.layers` and `DLWP.custom` resolve to this package's
    implementations.  Raises ImportError / AttributeError like the reference for unknown names.
    """
    mod = import_module(_MODULE_ALIASES.get(module_name, module_name))
    return getattr(mod, class_name)


def get_classes(module_name):
    """DLWP/util.py:96-108."""
    module = import_module(_MODULE_ALIASES.get(module_name, module_name))
    return {k: getattr(module, k) for k in dir(module) if isinstance(getattr(module, k), type)}


def get_methods(module_name):
    """DLWP/util.py:111-123."""
    module = import_module(_MODULE_ALIASES.get(module_name, module_name))
    return {k: getattr(module, k) for k in dir(module) if callable(getattr(module, k))}


def make_keras_picklable():
    """DLWP/util.py:27-49 patches keras.Model for pickling through a temp HDF5; this front end pickles natively."""
    return None


def save_model(model, file_name, history=None):
    """
    DLWP/util.py:126-153: `<file_name>.keras` (network + weights), `<file_name>.pkl` (wrapper with the network stripped)
    and optionally `<file_name>.history`.
    """
    if hasattr(model, 'base_model') and model.base_model is not None:
        model.base_model.save('%s.keras' % file_name)
    else:
        model.model.save('%s.keras' % file_name)
    model_copy = copy(model)
    model_copy.model = None
    if hasattr(model, 'base_model'):
        model_copy.base_model = None
    with open('%s.pkl' % file_name, 'wb') as f:
        pickle.dump(model_copy, f, protocol=pickle.HIGHEST_PROTOCOL)
    if history is not None:
        with open('%s.history' % file_name, 'wb') as f:
            pickle.dump(history.history, f, protocol=pickle.HIGHEST_PROTOCOL)


def load_model(file_name, history=False, custom_objects=None, gpus=1):
    """DLWP/util.py:156-192."""
    from .keras.models import load_model as _load
    with open('%s.pkl' % file_name, 'rb') as f:
        model = pickle.load(f)
    loaded = _load('%s.keras' % file_name, custom_objects=custom_objects, compile=True)
    model.base_model = loaded
    if gpus > 1:          # DLWP/util.py:176-183
        from .keras.utils import multi_gpu_model
        model.model = multi_gpu_model(loaded, gpus=gpus)
    else:
        model.model = loaded
    model.gpus = gpus
    if history:
        with open('%s.history' % file_name, 'rb') as f:
            h = pickle.load(f)
        return model, h
    return model


def day_of_year(date):
    """DLWP/util.py:300-302; also takes arrays of numpy datetime64."""
    d = np.asarray(date, 'datetime64[s]')
    start = d.astype('datetime64[Y]').astype('datetime64[s]')
    return (d - start).astype(np.float64) / 3600. / 24.


def insolation(dates, lat, lon, S=1.):
    """
    DLWP/util.py:305-352: approximate solar insolation for the given dates, (date, lat, lon) float32.  `dates`: anything
    numpy can read as datetime64 (pandas Timestamps included); `lat` / `lon` both 1-D or both 2-D of equal shape.  Unlike the
    reference the caller's `lat` array is not converted to radians in place.
    """
    lat, lon = np.asarray(lat, np.float64), np.asarray(lon, np.float64)
    if lat.ndim != lon.ndim:
        raise ValueError("'lat' and 'lon' must either both be 1d or both be 2d'")
    if lat.ndim == 2 and lat.shape != lon.shape:
        raise ValueError("shape mismatch between lat (%s) and lon (%s)" % (lat.shape, lon.shape))
    if lat.ndim == 1:
        lon, lat = np.meshgrid(lon, lat)
    # constants for year 1995 (standard)
    eps = 23.4441 * np.pi / 180.
    ecc = 0.016715
    om = 282.7 * np.pi / 180.
    beta = np.sqrt(1 - ecc ** 2.)
    days = day_of_year(dates).reshape(-1)
    lambda_m0 = ecc * (1. + beta) * np.sin(om)
    lambda_m = lambda_m0 + 2. * np.pi * (days - 80.5) / 365.
    lambda_ = lambda_m + 2. * ecc * np.sin(lambda_m - om)
    dec = np.arcsin(np.sin(eps) * np.sin(lambda_))                       # solar declination
    h = 2 * np.pi * (days[:, None, None] + lon / 360.)                   # hour angle
    rho = (1. - ecc ** 2.) / (1. + ecc * np.cos(lambda_ - om))           # distance
    lat = lat * np.pi / 180.
    sol = S * (np.sin(lat[None, ...]) * np.sin(dec[:, None, None]) -
               np.cos(lat[None, ...]) * np.cos(dec[:, None, None]) * np.cos(h)) * rho[:, None, None] ** -2.
    sol[sol < 0.] = 0.
    return sol.astype(np.float32)


def train_test_split_ind(n_sample, test_size, method='random'):
    """Index lists (train, test) splitting range(n_sample): same contract as DLWP/util.py:271-297."""
    idx = np.arange(n_sample)
    if method == 'first':
        test = idx[:test_size]
    elif method == 'last':
        test = idx[n_sample - test_size:]
    elif method == 'random':
        test = np.sort(np.random.permutation(n_sample)[:test_size])
    else:
        raise ValueError("'method' must be 'first', 'last', or 'random'")
    train = np.setdiff1d(idx, test)
    return train.tolist(), test.tolist()
