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


class ArraySeriesGenerator(object):
    """
    In-memory stand-in for the reference's `SeriesDataGenerator` (DLWP/model/generators.py:529-605) as far as
    `TimeSeriesEstimator` reads it (extensions.py:37-135, 175-188): unscaled predictors for every sample plus the metadata
    the estimator needs -- sample times, variable/level names of inputs and outputs, time steps, insolation flag.

    data: (n_sample + padding, V, H, W) time series of fields on an evenly spaced time axis `times`; sample i takes
    `input_time_steps` consecutive fields starting at i (selected `input_varlev`) and targets the `output_time_steps`
    fields that follow after `interval - 1` skipped steps (selected `output_varlev`).
    """

    def __init__(self, data, times, lat, lon, varlev, input_varlev=None, output_varlev=None, input_time_steps=1,
                 output_time_steps=1, interval=1, add_insolation=False):
        self.data = np.asarray(data, np.float32)
        self.times = np.asarray(times).astype('datetime64[s]')
        self.lat, self.lon = np.asarray(lat, np.float64), np.asarray(lon, np.float64)
        self.varlev = list(varlev)
        self._input_sel = {'varlev': list(input_varlev if input_varlev is not None else varlev)}
        self._output_sel = {'varlev': list(output_varlev if output_varlev is not None else varlev)}
        self._input_time_steps, self._output_time_steps = int(input_time_steps), int(output_time_steps)
        self._interval = int(interval)
        self._add_insolation = bool(add_insolation)
        self._n_sample = self.data.shape[0] - self._input_time_steps - self._output_time_steps - self._interval + 2
        if self._n_sample < 2:
            raise ValueError('not enough time steps for two samples')
        self.sample_times = self.times[:self._n_sample]
        H, W = self.data.shape[-2:]
        v_in = len(self._input_sel['varlev']) + (1 if self._add_insolation else 0)
        self.convolution_shape = (self._input_time_steps * v_in, H, W)
        self.output_convolution_shape = (self._output_time_steps * len(self._output_sel['varlev']), H, W)

    def generate(self, samples=(), scale_and_impute=False):
        from ..util import insolation
        S, ti, to = self._n_sample, self._input_time_steps, self._output_time_steps
        iin = [self.varlev.index(v) for v in self._input_sel['varlev']]
        iout = [self.varlev.index(v) for v in self._output_sel['varlev']]
        dt = self.times[1] - self.times[0]
        p = np.stack([self.data[n:n + S][:, iin] for n in range(ti)], axis=1)            # (S, ti, V_in, H, W)
        if self._add_insolation:
            sol = np.stack([insolation(self.sample_times + n * dt, self.lat, self.lon) for n in range(ti)], axis=1)
            p = np.concatenate([p, sol[:, :, None]], axis=2)
        off = ti + self._interval - 1
        t = np.stack([self.data[off + n:off + n + S][:, iout] for n in range(to)], axis=1)
        return p.reshape((S,) + self.convolution_shape), t.reshape((S,) + self.output_convolution_shape)


class DeviceSeriesGenerator(Sequence):
    """
    `SeriesDataGenerator` (DLWP/model/generators.py:425-640) with the data array RESIDENT ON THE GPU: `__getitem__` assembles
    a batch -- time-shifted predictor slices, the insolation channel, one target array per element of the sequence -- with
    `dlwp_gather_series` straight out of HBM and returns CUDA tensors, which `fit_generator` feeds to the training step
    without touching the host (the reference gathers with numpy fancy indexing from an xarray-backed array and Keras copies
    every batch to the device).  Same metadata as `ArraySeriesGenerator` (which it wraps); `sequence` = number of target
    arrays (examples/train_functional.py: integration_steps), batches shuffled per epoch like generators.py:146-159.
    """

    def __init__(self, series, batch_size=32, sequence=None, shuffle=False, seed=0):
        import ctypes
        import torch
        from .. import _native as nat
        self._nat, self._torch, self._ctypes = nat, torch, ctypes
        g = self.series = series
        self.batch_size, self.sequence, self.shuffle = int(batch_size), sequence, bool(shuffle)
        self._rng = np.random.RandomState(seed)
        nseq = int(sequence or 1)
        self._n_sample = g.data.shape[0] - g._input_time_steps - g._output_time_steps * nseq - g._interval + 2
        if self._n_sample < 1:
            raise ValueError('not enough time steps for one sample')
        self.data = torch.from_numpy(np.ascontiguousarray(g.data)).cuda()
        self._sel_in = torch.tensor([g.varlev.index(v) for v in g._input_sel['varlev']], dtype=torch.int32, device='cuda')
        self._sel_out = torch.tensor([g.varlev.index(v) for v in g._output_sel['varlev']], dtype=torch.int32, device='cuda')
        self._zero = torch.zeros(1, dtype=torch.int32, device='cuda')
        self.sol = None
        if g._add_insolation:
            from ..util import insolation
            self.sol = torch.from_numpy(insolation(g.times, g.lat, g.lon)).cuda()      # (n_times, H, W)
        self.convolution_shape = g.convolution_shape
        self.output_convolution_shape = g.output_convolution_shape
        self._idx = np.arange(self._n_sample)
        self.on_epoch_end()

    def __len__(self):
        return int(np.ceil(self._n_sample / self.batch_size))

    def on_epoch_end(self):
        if self.shuffle:
            self._rng.shuffle(self._idx)

    def _gather(self, data, samples, sel, out, T, V, V_total, t_off, c_per_t, c0):
        torch, nat = self._torch, self._nat
        H, W = data.shape[-2:]
        nat.check(nat.lib().dlwp_gather_series(data.data_ptr(), samples.data_ptr(), sel.data_ptr(), out.data_ptr(),
                                               samples.numel(), T, V, V_total, H, W, t_off, c_per_t, c0,
                                               self._ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  'dlwp_gather_series')

    def __getitem__(self, index):
        torch, g = self._torch, self.series
        sel = self._idx[index * self.batch_size:(index + 1) * self.batch_size]
        samples = torch.from_numpy(np.ascontiguousarray(sel, np.int64)).cuda()
        B, ti, to = len(sel), g._input_time_steps, g._output_time_steps
        H, W = g.data.shape[-2:]
        vi, vo, vt = self._sel_in.numel(), self._sel_out.numel(), g.data.shape[1]
        c_in = vi + (1 if self.sol is not None else 0)
        X = torch.empty((B, ti * c_in, H, W), dtype=torch.float32, device='cuda')
        self._gather(self.data, samples, self._sel_in, X, ti, vi, vt, 0, c_in, 0)
        if self.sol is not None:
            self._gather(self.sol, samples, self._zero, X, ti, 1, 1, 0, c_in, vi)
        ys = []
        for s in range(int(self.sequence or 1)):
            y = torch.empty((B, to * vo, H, W), dtype=torch.float32, device='cuda')
            self._gather(self.data, samples, self._sel_out, y, to, vo, vt, ti + g._interval - 1 + to * s, vo, 0)
            ys.append(y)
        return X, (ys if self.sequence is not None else ys[0])
