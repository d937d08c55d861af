"""
Names of `DLWP.model.generators` (reference DLWP/model/generators.py).  The reference's generators read xarray
Datasets from netCDF -- disk-bound host code outside the rollout hot path and xarray is not
available offline.  The classes exist so `isinstance` checks in `fit_generator` (models.py:224) and imports keep working;
`ArrayDataGenerator` is the in-memory stand-in used by tests and benchmarks: the same Sequence protocol
(`__len__`, `__getitem__`, `on_epoch_end`, `convolution_shape`) over numpy arrays.
"""

import numpy as np

from ..keras.utils import Sequence


class DataGenerator(Sequence):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('DLWP.model.DataGenerator needs xarray/netCDF4 (not available offline); use '
                                  'dlwp_b200.model.ArrayDataGenerator for in-memory arrays')


class SmartDataGenerator(DataGenerator):
    pass


class SeriesDataGenerator(DataGenerator):
    pass


class ArrayDataGenerator(Sequence):
    """Batches of (X, y) or (X, [y_1..y_k]) from in-memory arrays, shuffled per epoch like generators.py:146-159."""

    def __init__(self, X, y, batch_size=32, shuffle=False, seed=0):
        self.X = np.asarray(X, np.float32)
        self.y = [np.asarray(v, np.float32) for v in y] if isinstance(y, (list, tuple)) else np.asarray(y, np.float32)
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self._rng = np.random.RandomState(seed)
        self._idx = np.arange(self.X.shape[0])
        self.on_epoch_end()

    @property
    def convolution_shape(self):
        return tuple(self.X.shape[1:])

    @property
    def output_convolution_shape(self):
        y = self.y[0] if isinstance(self.y, list) else self.y
        return tuple(y.shape[1:])

    def __len__(self):
        return int(np.ceil(self.X.shape[0] / self.batch_size))

    def __getitem__(self, index):
        sel = self._idx[index * self.batch_size:(index + 1) * self.batch_size]
        if isinstance(self.y, list):
            return self.X[sel], [v[sel] for v in self.y]
        return self.X[sel], self.y[sel]

    def on_epoch_end(self):
        if self.shuffle:
            self._rng.shuffle(self._idx)
