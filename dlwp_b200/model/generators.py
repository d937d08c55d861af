"""
`DLWP.model.generators` (reference DLWP/model/generators.py).

* `SeriesDataGenerator` -- the generator of the reference's example scripts (generators.py:323-640) -- works on any dataset
  object that answers the handful of xarray calls it needs (`ds.predictors`, `ds.dims`, `.isel(time_step=-1)`,
  `.sel(**selection)`, `.values`, the `sample` / `lat` / `lon` coordinates): a real `xarray.Dataset` where xarray is
  installed (it is not in this image nor on the GPU box) or the numpy stand-in the golden generator uses.  Batches are
  assembled in memory exactly like the reference's `generate`; `as_array_series()` / `to_device()` hand the same data to the
  GPU-resident assembly (`DeviceSeriesGenerator`, `dlwp_gather_series`).  Pinned against the reference's own class run on
  the stand-in (tests/golden/series_generator.npz).
* `DataGenerator` / `SmartDataGenerator` (the latter deprecated in the reference itself) read netCDF-backed predictor /
  target pairs -- disk-bound host code outside the rollout path; the names exist for `isinstance`
  checks and imports and raise on construction.
* `ArrayDataGenerator` / `ArraySeriesGenerator`: in-memory stand-ins over numpy arrays (tests, benchmarks, TimeSeriesEstimator).
"""

import numpy as np

from ..keras.utils import Sequence


class DataGenerator(Sequence):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('DLWP.model.%s reads netCDF-backed predictor/target pairs (outside the rollout path); use '
                                  'SeriesDataGenerator, or dlwp_b200.model.ArrayDataGenerator for in-memory arrays'
                                  % type(self).__name__)


class SmartDataGenerator(DataGenerator):
    pass


def _drop_nan_samples(p, t):
    """DLWP/util.py:238-268 with the defaults: samples (rows) with a NaN in the predictors or the targets are removed."""
    bad = np.isnan(p.reshape(p.shape[0], -1)).any(axis=1) | np.isnan(t.reshape(t.shape[0], -1)).any(axis=1)
    return (p[~bad], t[~bad]) if bad.any() else (p, t)


class SeriesDataGenerator(Sequence):
    """
    generators.py:323-640: samples are windows of ONE continuous series -- `input_time_steps` consecutive fields as
    predictors, the `output_time_steps` fields that follow `interval - 1` skipped steps as targets (`sequence` such target
    arrays in a row for models that predict a sequence), an optional insolation channel per input time step.  Same
    constructor arguments, attributes (`_n_sample`, `_input_sel`, `_add_insolation`, ...), shape properties and Sequence
    protocol as the reference; `load` is accepted and ignored (the selected variables are held as numpy arrays).
    """

    def __init__(self, model, ds, input_sel=None, output_sel=None, input_time_steps=1, output_time_steps=1, sequence=None,
                 interval=1, add_insolation=False, batch_size=32, shuffle=False, remove_nan=True, load='required'):
        self.model = model
        if not hasattr(ds, 'predictors'):
            raise ValueError("dataset must have 'predictors' variable")
        for name, value in (('input_time_steps', input_time_steps), ('output_time_steps', output_time_steps),
                            ('batch_size', batch_size), ('interval', interval)):
            assert int(value) > 0, name
        if sequence is not None:
            assert int(sequence) > 0
        if load and load not in ('full', 'required', 'minimal') and not isinstance(load, bool):
            raise ValueError("'load' must be one of 'full', 'required', or 'minimal'")
        self.ds = ds
        self._batch_size, self._shuffle, self._remove_nan = batch_size, shuffle, remove_nan
        self._is_convolutional = model.is_convolutional
        self._keep_time_axis = model.is_recurrent
        self._impute_missing = model.impute
        self._sequence = sequence
        self._input_time_steps, self._output_time_steps, self._interval = input_time_steps, output_time_steps, interval
        self._n_sample = ds.dims['sample'] - input_time_steps - output_time_steps * (sequence or 1) + 2 - interval
        # ('time_step' datasets: the sample coordinate dates the LAST time step, generators.py:391-394)
        self.da = ds.predictors.isel(time_step=-1) if 'time_step' in ds.dims else ds.predictors
        self._input_sel, self._output_sel = input_sel or {}, output_sel or {}
        self.input_da = self.da.sel(**self._input_sel)
        self.output_da = self.da.sel(**self._output_sel)
        self._in = np.asarray(self.input_da.values)
        self._out = np.asarray(self.output_da.values)
        self._add_insolation = int(add_insolation)
        self._sol = None
        if add_insolation:
            from ..util import insolation
            self._sol = insolation(np.asarray(self.da.sample.values), np.asarray(self.da.lat.values, np.float64),
                                   np.asarray(self.da.lon.values, np.float64))
        self._indices = []
        self.on_epoch_end()

    # -- metadata the TimeSeriesEstimator of this package reads ------------------------------------------------------------------
    @property
    def sample_times(self):
        return np.asarray(self.ds.sample.values).astype('datetime64[s]')[:self._n_sample]

    @property
    def lat(self):
        return np.asarray(self.ds.lat.values, np.float64)

    @property
    def lon(self):
        return np.asarray(self.ds.lon.values, np.float64)

    # -- shapes (generators.py:425-522) ----------------------------------------------------------------------------------------
    def _shapes(self, steps, field_shape, sol):
        """(original, features, dense, convolution) shapes of `steps` stacked fields of shape `field_shape`."""
        full = (steps,) + tuple(field_shape)
        grid = full[-2:]
        n = int(np.prod(full)) + int(np.prod(grid)) * steps * sol
        dense = (steps, n // steps) if self._keep_time_axis else (n,)
        if self._keep_time_axis:
            conv = (steps, int(np.prod(full[1:-2])) + sol) + grid
        else:
            conv = (int(np.prod(full[:-2])) + steps * sol,) + grid
        return full, n, dense, conv

    @property
    def shape(self):
        return self._shapes(self._input_time_steps, self._in.shape[1:], 0)[0]

    @property
    def n_features(self):
        return self._shapes(self._input_time_steps, self._in.shape[1:], self._add_insolation)[1]

    @property
    def dense_shape(self):
        return self._shapes(self._input_time_steps, self._in.shape[1:], self._add_insolation)[2]

    @property
    def convolution_shape(self):
        return self._shapes(self._input_time_steps, self._in.shape[1:], self._add_insolation)[3]

    @property
    def output_shape(self):
        return self._shapes(self._output_time_steps, self._out.shape[1:], 0)[0]

    @property
    def output_n_features(self):
        return self._shapes(self._output_time_steps, self._out.shape[1:], 0)[1]

    @property
    def output_dense_shape(self):
        return self._shapes(self._output_time_steps, self._out.shape[1:], 0)[2]

    @property
    def output_convolution_shape(self):
        return self._shapes(self._output_time_steps, self._out.shape[1:], 0)[3]

    def _as_2d(self, which):
        keep, self._keep_time_axis = self._keep_time_axis, False
        try:
            return tuple(getattr(self, which))
        finally:
            self._keep_time_axis = keep

    @property
    def shape_2d(self):
        return self._as_2d('convolution_shape')

    @property
    def output_shape_2d(self):
        return self._as_2d('output_convolution_shape')

    # -- batches (generators.py:524-640) -----------------------------------------------------------------------------------------
    def on_epoch_end(self):
        self._indices = np.arange(self._n_sample)
        if self._shuffle:
            np.random.shuffle(self._indices)

    @staticmethod
    def _windows(series, samples, first, steps):
        """(len(samples), steps, ...): `steps` consecutive fields starting `first` after each sample index."""
        return np.stack([series[samples + first + n] for n in range(steps)], axis=1)

    def _finish(self, p, t, scale_and_impute):
        if self._remove_nan:
            p, t = _drop_nan_samples(p, t)
        if scale_and_impute:
            if self._impute_missing:
                p, t = self.model.imputer_transform(p, t)
            p, t = self.model.scaler_transform(p, t)
        if self._is_convolutional:
            p = p.reshape((p.shape[0],) + tuple(self.convolution_shape))
            t = t.reshape((t.shape[0],) + tuple(self.output_convolution_shape))
        elif self._keep_time_axis:
            p = p.reshape((p.shape[0],) + tuple(self.dense_shape))
            t = t.reshape((t.shape[0],) + tuple(self.output_dense_shape))
        return p, t

    def generate(self, samples, scale_and_impute=True):
        samples = np.arange(self._n_sample) if len(samples) == 0 else np.array(samples, dtype=np.int64)
        ti, to = self._input_time_steps, self._output_time_steps
        p = self._windows(self._in, samples, 0, ti)
        if self._sol is not None:                      # the insolation channel goes last within every input time step
            sol = self._windows(self._sol, samples, 0, ti)
            p = np.concatenate([p.reshape((len(samples), ti, -1) + sol.shape[-2:]), sol[:, :, None]], axis=2)
        p = p.reshape((len(samples), -1))
        first = ti + self._interval - 1
        if self._sequence is None:
            t = self._windows(self._out, samples, first, to).reshape((len(samples), -1))
            return self._finish(p, t, scale_and_impute)
        targets = []
        for s in range(self._sequence):                # (like the reference, the predictors pass through _finish every time)
            t = self._windows(self._out, samples, first + to * s, to).reshape((len(samples), -1))
            p, t = self._finish(p.reshape((p.shape[0], -1)), t, scale_and_impute)
            targets.append(t)
        return p, targets

    def __len__(self):
        return int(np.ceil(self._n_sample / self._batch_size))

    def __getitem__(self, index):
        if int(index) < 0:
            index = len(self) + index
        if index > len(self):
            raise IndexError
        return self.generate(self._indices[index * self._batch_size:(index + 1) * self._batch_size])

    # -- hand-over to the GPU-resident assembly --------------------------------------------------------------------------------
    def as_array_series(self):
        """The same series as an `ArraySeriesGenerator` ('varlev' datasets: fields of shape (varlev, lat, lon))."""
        if 'varlev' not in self.ds.dims or self._in.ndim != 4 or set(self._input_sel) - {'varlev'} or \
                set(self._output_sel) - {'varlev'}:
            raise NotImplementedError('only datasets with a flat varlev dimension map onto ArraySeriesGenerator')
        names = [str(v) for v in np.asarray(self.da.varlev.values)]
        return ArraySeriesGenerator(np.asarray(self.da.values, np.float32), np.asarray(self.da.sample.values), self.lat,
                                    self.lon, names, [str(v) for v in self._input_sel.get('varlev', names)],
                                    [str(v) for v in self._output_sel.get('varlev', names)], self._input_time_steps,
                                    self._output_time_steps, self._interval, bool(self._add_insolation))

    def to_device(self, seed=0):
        """`DeviceSeriesGenerator` over the same series: batches gathered on the GPU, CUDA tensors for fit_generator."""
        return DeviceSeriesGenerator(self.as_array_series(), batch_size=self._batch_size, sequence=self._sequence,
                                     shuffle=self._shuffle, seed=seed)


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
