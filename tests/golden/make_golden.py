"""
Generate the golden fixtures in this directory by running the REFERENCE'S OWN CODE in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only); writes tests/golden/*.npz

The reference (jweyn/DLWP @ 3f32bfab) cannot be imported as a package here: ``DLWP/model/__init__.py`` pulls in keras,
tensorflow, xarray and netCDF4, none of which are installed or installable offline.  Its hot-path modules are therefore
loaded file by file with stub ``keras`` / ``tensorflow`` modules in ``sys.modules``:

* ``DLWP/custom.py``            -> the real ``PeriodicPadding2D.call`` (custom.py:191-214) and ``row_conv2d`` (840-896)
                                   run on numpy arrays through a numpy-backed stub of ``keras.backend``;
* ``DLWP/model/models.py``      -> the real ``DLWPNeuralNet`` / ``DLWPFunctional`` ``predict`` + ``predict_timeseries``
                                   (230-301, 404-452) run with a duck-typed ``.model``;
* ``DLWP/model/models_torch.py``-> the real ``DLWPTorchNN`` (77-376) runs end to end on torch-CPU.

Nothing from the reference is copied: only inputs, weights and the arrays the reference code returned are stored.
This script is never run on the GPU box (``/root/reference`` does not exist there); the committed ``.npz`` files are
what the tests read.
"""

import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('DLWP_REFERENCE', '/root/reference')
sys.path.insert(0, REPO)

from oracle import layers as OL  # noqa: E402
from oracle import ops as OO  # noqa: E402


# ------------------------------------------------------------------------------------------------------------------ #
# Stub third-party modules the reference imports at module level
# ------------------------------------------------------------------------------------------------------------------ #

def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)  # torch._dynamo probes find_spec('tensorflow')
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _StubZeroPadding2D(object):
    """Just enough of keras.layers.ZeroPadding2D.__init__ (argument normalisation) for PeriodicPadding2D to subclass."""

    def __init__(self, padding=(1, 1), data_format=None, **kwargs):
        self.data_format = 'channels_last' if data_format is None else data_format
        if isinstance(padding, int):
            self.padding = ((padding, padding), (padding, padding))
        else:
            pads = []
            for p in padding:
                pads.append((p, p) if isinstance(p, int) else tuple(p))
            self.padding = tuple(pads)


class _StubZeroPadding3D(_StubZeroPadding2D):
    """keras.layers.ZeroPadding3D.__init__ argument normalisation (int -> three symmetric pairs)."""

    def __init__(self, padding=(1, 1, 1), data_format=None, **kwargs):
        super(_StubZeroPadding3D, self).__init__((padding,) * 3 if isinstance(padding, int) else padding, data_format)


class _Assignable(np.ndarray):
    """K.zeros(...) with an eager `.assign` (see install_stubs)."""

    def __new__(cls, shape):
        return np.zeros(shape).view(cls)

    def assign(self, value):
        self[...] = value
        return self


class _Any(object):
    def __init__(self, *a, **k):
        pass


def _np_conv2d(x, kernel, strides=(1, 1), padding='valid', data_format=None):
    assert padding == 'valid'
    return OO.conv2d_valid(x, kernel, None, (1, 1), strides, data_format)


def install_stubs():
    K = _mod('keras.backend',
             backend=lambda: 'numpy',
             concatenate=lambda xs, axis=-1: np.concatenate(xs, axis=axis),
             stack=lambda xs, axis=0: np.stack(xs, axis=axis),
             normalize_data_format=lambda v: 'channels_last' if v is None else v,
             conv2d=_np_conv2d,
             # the reductions / pointwise ops the loss functions of custom.py:994-1088 use
             mean=lambda x, axis=None: np.mean(x, axis=tuple(axis) if isinstance(axis, list) else axis),
             sqrt=np.sqrt, square=np.square, abs=np.abs, variable=lambda v, name=None: np.asarray(v),
             # latitude_weighted_loss (custom.py:956-991); `zeros` returns an array whose `assign` stores eagerly (TF2 / the
             # documented intent -- under graph-mode TF1 the assign op of custom.py:973 is never run)
             cos=np.cos, sin=np.sin, pow=np.power, cast_to_floatx=lambda v: np.asarray(v, np.float32),
             expand_dims=lambda x, axis=-1: np.expand_dims(x, axis),
             repeat_elements=lambda x, rep, axis: np.repeat(x, rep, axis=axis))
    K.zeros = lambda shape, **kw: _Assignable(shape)
    K.ones = np.ones
    _mod('keras', backend=K)
    _mod('keras.callbacks', Callback=_Any, EarlyStopping=_Any)
    _mod('keras.layers', Lambda=_Any, Layer=_Any)
    _mod('keras.layers.convolutional', ZeroPadding2D=_StubZeroPadding2D, ZeroPadding3D=_StubZeroPadding3D)
    _mod('keras.layers.local', LocallyConnected2D=_Any)
    # keras.losses (Keras 2.2 losses.py): the mean over the LAST axis
    _mod('keras.losses', mean_absolute_error=lambda y_true, y_pred: np.mean(np.abs(y_pred - y_true), axis=-1),
         mean_squared_error=lambda y_true, y_pred: np.mean(np.square(y_pred - y_true), axis=-1))
    _mod('keras.utils', conv_utils=None, multi_gpu_model=None, Sequence=object)
    _mod('keras.engine')
    _mod('keras.engine.base_layer', InputSpec=_Any)
    _mod('keras.models')
    _mod('tensorflow')
    sys.modules['keras'].layers = sys.modules['keras.layers']
    sys.modules['keras'].models = sys.modules['keras.models']
    # parent packages for the relative imports inside the reference files
    pkg = _mod('DLWP')
    pkg.__path__ = []
    mpkg = _mod('DLWP.model')
    mpkg.__path__ = []
    _mod('DLWP.model.generators', DataGenerator=type('DataGenerator', (), {}),
         SmartDataGenerator=type('SmartDataGenerator', (), {}),
         SeriesDataGenerator=type('SeriesDataGenerator', (), {}))


def load_ref(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------------------------ #

def gen_periodic_padding(custom):
    rng = np.random.RandomState(10)
    out = {}
    x_cf = rng.standard_normal((2, 3, 5, 7)).astype(np.float32)
    x_cl = rng.standard_normal((2, 5, 7, 3)).astype(np.float32)
    out['x_channels_first'] = x_cf
    out['x_channels_last'] = x_cl
    pads = [(0, 2), (1, 1), ((1, 2), (3, 0)), 2, ((0, 0), (2, 2)), (2, 0), ((0, 1), (0, 3))]
    out['n_cases'] = np.int64(len(pads))
    for k, p in enumerate(pads):
        out['pad_%d' % k] = np.asarray(OO.normalize_padding(p), np.int64)
        for fmt, x in (('channels_first', x_cf), ('channels_last', x_cl)):
            layer = custom.PeriodicPadding2D(padding=p, data_format=fmt)
            out['y_%d_%s' % (k, fmt)] = layer.call(x)
    np.savez_compressed(os.path.join(HERE, 'periodic_padding2d.npz'), **out)


def gen_padding_3d_and_fill(custom):
    """The real PeriodicPadding3D.call (custom.py:277-306) and FillPadding2D.call (custom.py:359-402) on numpy arrays."""
    rng = np.random.RandomState(15)
    out = {}
    x5_cf = rng.standard_normal((2, 3, 4, 5, 7)).astype(np.float32)   # (batch, depth, a1, a2, a3)
    x5_cl = rng.standard_normal((2, 4, 5, 7, 3)).astype(np.float32)
    out['x5_channels_first'], out['x5_channels_last'] = x5_cf, x5_cl
    pads3 = [(0, 0, 2), (0, 2, 0), (1, 1, 1), ((0, 1), (2, 0), (1, 3)), 2]
    out['n3'] = np.int64(len(pads3))
    for k, p in enumerate(pads3):
        out['pad3_%d' % k] = np.asarray(OO.normalize_padding3d(p), np.int64)
        for fmt, x in (('channels_first', x5_cf), ('channels_last', x5_cl)):
            out['y3_%d_%s' % (k, fmt)] = custom.PeriodicPadding3D(padding=p, data_format=fmt).call(x)
    x_cf = rng.standard_normal((2, 3, 5, 7)).astype(np.float32)
    x_cl = rng.standard_normal((2, 5, 7, 3)).astype(np.float32)
    out['x_channels_first'], out['x_channels_last'] = x_cf, x_cl
    pads2 = [(0, 2), (1, 1), ((1, 2), (3, 0)), 2, ((0, 0), (2, 2)), ((0, 1), (0, 3))]
    out['n2'] = np.int64(len(pads2))
    for k, p in enumerate(pads2):
        out['pad2_%d' % k] = np.asarray(OO.normalize_padding(p), np.int64)
        for fmt, x in (('channels_first', x_cf), ('channels_last', x_cl)):
            out['yfill_%d_%s' % (k, fmt)] = custom.FillPadding2D(padding=p, data_format=fmt).call(x)
    np.savez_compressed(os.path.join(HERE, 'padding3d_fill2d.npz'), **out)


def _small_recurrent(time_dim, nvar, H, W, seed):
    """The recurrent front block of examples/train.py:144-157 in front of a small conv stack."""
    cf = 'channels_first'
    cs = (time_dim, nvar, H, W)
    net = OL.OSequential((
        ('PeriodicPadding3D', ((0, 0, 2),), {'data_format': cf, 'input_shape': cs}),
        ('ZeroPadding3D', ((0, 2, 0),), {'data_format': cf}),
        ('ConvLSTM2D', (2 * nvar, 3), {'dilation_rate': 2, 'padding': 'valid', 'data_format': cf, 'activation': 'tanh',
                                       'return_sequences': True}),
        ('Reshape', ((2 * time_dim * nvar, H, W),), None),
        ('PeriodicPadding2D', ((0, 1),), {'data_format': cf}),
        ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
        ('Conv2D', (time_dim * nvar, 3), {'activation': 'linear', 'data_format': cf}),
        ('Reshape', (cs,), None),
    ))
    rng = np.random.RandomState(seed)
    for layer in net.weight_layers:
        if isinstance(layer, OL.OConvLSTM2D):
            layer.randomize(rng, bias_scale=0.1)
    OL.init_weights(net.conv_layers, seed=seed + 1, bias_scale=0.1)
    return net


def gen_rollout_recurrent(models):
    """The reference's own DLWPNeuralNet.predict_timeseries with is_recurrent=True (models.py:270-301: 5-D predictors,
    feature_shape = shape[2:]) around a ConvLSTM2D-fronted net; the net's arithmetic is the oracle's."""
    out = {}
    rng = np.random.RandomState(16)
    cases = []
    for time_dim in (2, 3):
        net = _small_recurrent(time_dim, 2, 6, 8, seed=40 + time_dim)
        x0 = rng.standard_normal((3, time_dim, 2, 6, 8)).astype(np.float32)
        out['x0_td%d' % time_dim] = x0
        for k, w in enumerate(net.get_weights()):
            out['w_td%d_%d' % (time_dim, k)] = w
        dlwp = models.DLWPNeuralNet(is_convolutional=True, is_recurrent=True, time_dim=time_dim, scaler_type=None,
                                    scale_targets=False)
        dlwp.model = _FakeKerasModel(net)
        for steps in (1, 5):
            for ss in (False, True):
                for ktd in (False, True):
                    key = 'y_td%d_s%d_ss%d_k%d' % (time_dim, steps, ss, ktd)
                    out[key] = dlwp.predict_timeseries(x0, steps, step_sequence=ss, keep_time_dim=ktd)
                    cases.append(key)
    out['cases'] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, 'rollout_recurrent.npz'), **out)


def gen_insolation():
    """The reference's own `day_of_year` / `insolation` (DLWP/util.py:300-352), executed from its source text (util.py
    imports keras at module level; these two functions only need numpy and pandas)."""
    import pandas as pd
    src = open(os.path.join(REF, 'DLWP', 'util.py')).read()
    a, b = src.index('def day_of_year'), src.index('return sol.astype(np.float32)') + len('return sol.astype(np.float32)')
    ns = {'np': np, 'pd': pd}
    exec(compile(src[a:b], 'DLWP/util.py', 'exec'), ns)
    dates = pd.date_range('2003-02-27 06:00', periods=11, freq='6h')
    lat, lon = np.linspace(90., -90., 13), np.arange(0., 360., 30.)
    sol = ns['insolation'](dates, lat.copy(), lon.copy())
    lon2, lat2 = np.meshgrid(lon, lat)
    sol2 = ns['insolation'](dates[:3], lat2.copy(), lon2.copy(), S=2.)
    np.savez_compressed(os.path.join(HERE, 'insolation.npz'), dates=dates.values.astype('datetime64[s]').astype(np.int64),
                        lat=lat, lon=lon, sol=sol, sol2=sol2)


def gen_acc_loss(custom):
    """DLWP/custom.py:1036-1088 anomaly_correlation_loss, every regularize_mean, with and without a climatology, and
    custom.py:994-1033 anomaly_correlation -- executed from the reference module on float64 arrays."""
    rng = np.random.RandomState(21)
    y_true = rng.standard_normal((3, 4, 6, 8)) + 0.5
    y_pred = y_true + 0.4 * rng.standard_normal((3, 4, 6, 8))
    mean = 0.3 * rng.standard_normal((1, 4, 6, 8))
    out = {'y_true': y_true, 'y_pred': y_pred, 'mean': mean}
    for reg in (None, 'mse', 'mae', 'global', 'spatial'):
        for use_mean in (False, True):
            fn = custom.anomaly_correlation_loss(mean=mean if use_mean else None, regularize_mean=reg)
            out['loss_%s_%d' % (reg, use_mean)] = np.asarray(fn(y_true, y_pred), np.float64)
        out['metric_%s' % reg] = np.asarray(custom.anomaly_correlation(y_true, y_pred, regularize_mean=reg), np.float64)
    out['loss_none_forward'] = np.asarray(custom.anomaly_correlation_loss(regularize_mean=None, reverse=False)(y_true, y_pred))
    np.savez_compressed(os.path.join(HERE, 'acc_loss.npz'), **out)


def gen_estimator(models, util):
    """The reference's OWN `TimeSeriesEstimator.__init__` / `.predict` (DLWP/model/extensions.py:21-303) executed on a
    numpy stand-in for xarray (tests/golden/fake_xarray.py: reindex = exact label lookup with NaN fill, .loc = label ->
    position) with a SeriesDataGenerator-shaped stub and a model whose `predict` is a fixed numpy map.  Cases: equal
    input / output time steps; fewer output than input steps with insolation and an output varlev subset; more output than
    input steps with prefer_first_times on and off; impute with interval 2; keep_time_dim."""
    import fake_xarray
    sys.modules['xarray'] = fake_xarray
    ext = load_ref('DLWP.model.extensions', 'DLWP/model/extensions.py')
    gens = sys.modules['DLWP.model.generators']
    rng = np.random.RandomState(33)
    names = np.array(['z/500', 't/850', 'u/300'])
    nt, H, W = 14, 5, 8
    data = rng.standard_normal((nt, 3, H, W)).astype(np.float32)
    times = np.datetime64('2003-03-01T00:00', 'ns') + np.arange(nt) * np.timedelta64(6 * 3600 * 10 ** 9, 'ns')
    lat, lon = np.linspace(80., -80., H), np.arange(0., 360., 45.)

    class Gen(gens.SeriesDataGenerator):
        def __init__(self, in_sel, out_sel, t_in, t_out, interval, sol):
            self.ds = fake_xarray.Dataset({'sample': times, 'varlev': names, 'lat': lat, 'lon': lon})
            self._input_sel = {'varlev': list(in_sel)} if in_sel is not None else {}
            self._output_sel = {'varlev': list(out_sel)} if out_sel is not None else {}
            self._input_time_steps, self._output_time_steps, self._interval = t_in, t_out, interval
            self._add_insolation = sol
            self._n_sample = nt - t_in - t_out - interval + 2
            iin = [list(names).index(v) for v in (in_sel if in_sel is not None else names)]
            iout = [list(names).index(v) for v in (out_sel if out_sel is not None else names)]
            S = self._n_sample
            pp = np.stack([data[n:n + S][:, iin] for n in range(t_in)], axis=1)
            if sol:
                dt = times[1] - times[0]
                so = np.stack([util.insolation(times[:S] + n * dt, lat.copy(), lon.copy()) for n in range(t_in)], axis=1)
                pp = np.concatenate([pp, so[:, :, None]], axis=2)
            off = t_in + interval - 1
            tt = np.stack([data[off + n:off + n + S][:, iout] for n in range(t_out)], axis=1)
            self.convolution_shape = (pp.shape[1] * pp.shape[2], H, W)
            self._p, self._t = pp.reshape((S, -1, H, W)), tt.reshape((S, -1, H, W))

        def generate(self, samples, scale_and_impute=True):
            return self._p.copy(), self._t.copy()

    class Net(models.DLWPNeuralNet):
        def __init__(self, w, time_dim):
            self.w, self.time_dim = w, time_dim

        def predict(self, x, **kwargs):
            return np.tanh(np.einsum('nchw,co->nohw', np.asarray(x, np.float32), self.w)).astype(np.float32)

    out = {'lat': lat, 'lon': lon, 'times': times.astype('datetime64[s]').astype(np.int64), 'names': names}
    cases = []
    specs = [('equal', None, None, 2, 2, 1, False, 5, {}),
             ('fewer_out_sol', ['z/500', 't/850', 'u/300'], ['z/500', 'u/300'], 2, 1, 1, True, 4, {}),
             ('more_out_first', None, None, 1, 2, 1, False, 3, {}),
             ('more_out_last', None, None, 1, 2, 1, False, 3, {'prefer_first_times': False}),
             ('impute_interval2', ['z/500', 't/850', 'u/300'], ['t/850'], 2, 1, 2, True, 3, {'impute': True}),
             ('equal_keep_time', None, None, 2, 2, 1, False, 5, {'keep_time_dim': True})]
    for key, in_sel, out_sel, t_in, t_out, interval, sol, steps, kw in specs:
        gen = Gen(in_sel, out_sel, t_in, t_out, interval, sol)
        v_in, v_out = gen._p.shape[1], gen._t.shape[1]
        w = (0.5 * rng.standard_normal((v_in, v_out))).astype(np.float32)
        est = ext.TimeSeriesEstimator(Net(w, t_in), gen)
        res = est.predict(steps, **kw)
        out[key + '/p'], out[key + '/t'], out[key + '/w'] = gen._p, gen._t, w
        out[key + '/result'] = np.asarray(res.values, np.float32)
        out[key + '/dims'] = np.array(res.dims)
        out[key + '/f_hour'] = np.asarray(res.coords['f_hour']).astype('timedelta64[s]').astype(np.int64)
        out[key + '/time'] = np.asarray(res.coords['time']).astype('datetime64[s]').astype(np.int64)
        out[key + '/varlev'] = np.asarray(res.coords['varlev'])
        out[key + '/spec'] = np.array([t_in, t_out, interval, int(sol), steps, int(kw.get('impute', False)),
                                       int(kw.get('prefer_first_times', True)), int(kw.get('keep_time_dim', False))])
        out[key + '/in_varlev'] = np.array(list(est._input_sel['varlev']))
        cases.append(key)
    # a DLWPFunctional that predicts a sequence (_n_steps = 2): the estimator hands the whole forecast to the model's own
    # predict_timeseries (extensions.py:204-208)
    class Chain(object):
        n_outputs = 2

        def __init__(self, w):
            self.w = w

        def forward(self, x):
            a = np.tanh(np.einsum('nchw,co->nohw', x, self.w))
            return [a, np.tanh(np.einsum('nchw,co->nohw', a, self.w))]

    gen = Gen(None, None, 1, 1, 1, False)
    w = (0.5 * rng.standard_normal((3, 3))).astype(np.float32)
    fun = models.DLWPFunctional(is_convolutional=True, is_recurrent=False, time_dim=1)
    fun.model = _FakeKerasModel(Chain(w.astype(np.float64)))
    fun._n_steps = 2
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        est = ext.TimeSeriesEstimator(fun, gen)
    res = est.predict(5)
    key = 'functional_sequence'
    out[key + '/p'], out[key + '/t'], out[key + '/w'] = gen._p, gen._t, w
    out[key + '/result'] = np.asarray(res.values, np.float32)
    out[key + '/dims'] = np.array(res.dims)
    out[key + '/f_hour'] = np.asarray(res.coords['f_hour']).astype('timedelta64[s]').astype(np.int64)
    out[key + '/time'] = np.asarray(res.coords['time']).astype('datetime64[s]').astype(np.int64)
    out['cases'] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, 'estimator.npz'), **out)


def gen_series_generator():
    """The reference's OWN `SeriesDataGenerator` (DLWP/model/generators.py:323-640: __init__, generate, __getitem__, __len__,
    the shape properties) on the xarray stand-in: predictors / targets of every sample and of one batch, with a variable
    selection, insolation, a target sequence and interval 2.  (`np.int`, which the reference still uses, is aliased to int:
    it left numpy in 1.24.)"""
    import fake_xarray
    sys.modules['xarray'] = fake_xarray
    if not hasattr(np, 'int'):
        np.int = int
    gmod = load_ref('DLWP.model.generators_impl', 'DLWP/model/generators.py')
    rng = np.random.RandomState(44)
    names = np.array(['z/500', 't/850', 'u/300', 'v/300'])
    nt, H, W = 15, 4, 6
    data = rng.standard_normal((nt, 4, H, W)).astype(np.float32)
    times = np.datetime64('2004-12-30T00:00', 'ns') + np.arange(nt) * np.timedelta64(6 * 3600 * 10 ** 9, 'ns')
    lat, lon = np.linspace(75., -75., H), np.arange(0., 360., 60.)
    pred = fake_xarray.DataArray(data, coords=[times, names, lat, lon], dims=['sample', 'varlev', 'lat', 'lon'])
    ds = fake_xarray.Dataset({'sample': times, 'varlev': names, 'lat': lat, 'lon': lon}, predictors=pred)

    class M(object):
        is_convolutional, is_recurrent, impute = True, False, False

        def scaler_transform(self, p, t):                 # scaler_type=None model: models.py scaler_transform is the identity
            return p, t

    out = {'data': data, 'times': times.astype('datetime64[s]').astype(np.int64), 'lat': lat, 'lon': lon, 'names': names}
    cases = []
    specs = [('plain', None, None, 1, 1, None, 1, False, 4),
             ('subset_sol', ['z/500', 'u/300', 't/850'], ['t/850', 'z/500'], 2, 1, None, 1, True, 4),
             ('sequence_interval2', ['z/500', 't/850', 'u/300', 'v/300'], ['u/300'], 2, 2, 3, 2, True, 3)]
    for key, in_sel, out_sel, t_in, t_out, seq, interval, sol, batch in specs:
        gen = gmod.SeriesDataGenerator(M(), ds, input_sel={'varlev': in_sel} if in_sel else None,
                                       output_sel={'varlev': out_sel} if out_sel else None, input_time_steps=t_in,
                                       output_time_steps=t_out, sequence=seq, interval=interval, add_insolation=sol,
                                       batch_size=batch, shuffle=False, remove_nan=False)
        p, t = gen.generate([], scale_and_impute=False)
        xb, yb = gen[1]                                   # second batch (scaler: identity model below)
        out[key + '/p'] = p
        out[key + '/xb'] = xb
        for k, (tt, yy) in enumerate(zip(t if seq else [t], yb if seq else [yb])):
            out[key + '/t%d' % k], out[key + '/yb%d' % k] = tt, yy
        out[key + '/spec'] = np.array([t_in, t_out, seq or 0, interval, int(sol), batch, gen._n_sample, len(gen)])
        out[key + '/shapes'] = np.array(list(gen.convolution_shape) + list(gen.output_convolution_shape))
        out[key + '/in_sel'] = np.array(in_sel if in_sel else list(names))
        out[key + '/out_sel'] = np.array(out_sel if out_sel else list(names))
        cases.append(key)
    out['cases'] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, 'series_generator.npz'), **out)


def gen_wrapper_pickles(models):
    """What `DLWP.util.save_model` writes as `<name>.pkl` (util.py:143-149): the reference's wrapper objects with `.model`
    and `.base_model` set to None, pickled by the reference's own classes -- one DLWPNeuralNet with fitted sklearn scalers,
    one DLWPFunctional that predicts a sequence."""
    import pickle
    from copy import copy
    rng = np.random.RandomState(8)
    nn = models.DLWPNeuralNet(is_convolutional=False, is_recurrent=False, time_dim=2, scaler_type='StandardScaler',
                              scale_targets=True, apply_same_y_scaling=False, impute_missing=False)
    X, y = rng.standard_normal((20, 6)), 3.0 * rng.standard_normal((20, 4)) + 1.0
    nn.scaler_fit(X, y)
    nn._is_init_fit = True
    fun = models.DLWPFunctional(is_convolutional=True, is_recurrent=False, time_dim=3)
    fun._n_steps, fun.gpus = 4, 2
    blobs = {}
    for key, obj in (('neuralnet', nn), ('functional', fun)):
        c = copy(obj)
        c.model = None
        c.base_model = None
        blobs[key] = np.frombuffer(pickle.dumps(c, protocol=pickle.HIGHEST_PROTOCOL), np.uint8)
    Xs, ys = nn.scaler_transform(X, y)
    np.savez_compressed(os.path.join(HERE, 'wrapper_pickles.npz'), X=X, y=y, Xs=Xs, ys=ys, **blobs)


def gen_lat_loss(custom):
    """latitude_weighted_loss (custom.py:956-991) with both weightings on a (C, H, W) output, and without latitudes."""
    rng = np.random.RandomState(17)
    lats = np.linspace(87.5, -87.5, 8)
    shape = (3, 8, 10)
    y_true = rng.standard_normal((4,) + shape).astype(np.float32)
    y_pred = (y_true + 0.3 * rng.standard_normal((4,) + shape)).astype(np.float32)
    out = {'lats': lats, 'y_true': y_true, 'y_pred': y_pred}
    mse = sys.modules['keras.losses'].mean_squared_error
    for weighting in ('cosine', 'midlatitude'):
        fn = custom.latitude_weighted_loss(mse, lats, shape, axis=-2, weighting=weighting)
        out['loss_' + weighting] = np.asarray(fn(y_true, y_pred), np.float64)
    out['loss_none'] = np.asarray(custom.latitude_weighted_loss(mse, None, shape)(y_true, y_pred), np.float64)
    np.savez_compressed(os.path.join(HERE, 'lat_loss.npz'), **out)


def gen_row_conv(custom):
    rng = np.random.RandomState(11)
    x = rng.standard_normal((2, 4, 9, 12)).astype(np.float64)
    kernel = rng.standard_normal((5, 5, 5, 4, 3)).astype(np.float64) * 0.2  # (H_out, kh, kw, Cin, Cout)
    y = custom.row_conv2d(x, kernel, (5, 5), (1, 1), (5, 8), 'channels_first')
    x_cl = np.moveaxis(x, 1, 3)
    y_cl = custom.row_conv2d(x_cl, kernel, (5, 5), (1, 1), (5, 8), 'channels_last')
    np.savez_compressed(os.path.join(HERE, 'row_conv2d.npz'), x=x, kernel=kernel, y=y, y_channels_last=y_cl)


class _FakeKerasModel(object):
    """Duck-typed stand-in for the compiled keras model: ``predict`` = oracle forward (float64 -> float32)."""

    def __init__(self, net):
        self.net = net
        self.outputs = [None] * net.n_outputs

    def predict(self, x, **kwargs):
        y = self.net.forward(np.asarray(x, np.float64))
        if isinstance(y, list):
            return [v.astype(np.float32) for v in y]
        return y.astype(np.float32)


def _small_seq(time_dim, nvar, H, W, seed):
    cf = 'channels_first'
    C = time_dim * nvar
    net = OL.OSequential((
        ('PeriodicPadding2D', ((0, 1),), {'data_format': cf, 'input_shape': (C, H, W)}),
        ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
        ('Conv2D', (8, 3), {'activation': 'tanh', 'data_format': cf}),
        ('PeriodicPadding2D', ((0, 2),), {'data_format': cf}),
        ('ZeroPadding2D', ((2, 0),), {'data_format': cf}),
        ('Conv2D', (C, 3), {'dilation_rate': 2, 'activation': 'linear', 'data_format': cf}),
    ))
    OL.init_weights(net.conv_layers, seed=seed, bias_scale=0.1)
    return net


def gen_rollout_neuralnet(models):
    out = {}
    rng = np.random.RandomState(12)
    cases = []
    for time_dim in (1, 2, 3):
        net = _small_seq(time_dim, 2, 6, 8, seed=20 + time_dim)
        x0 = rng.standard_normal((3, time_dim * 2, 6, 8)).astype(np.float32)
        out['x0_td%d' % time_dim] = x0
        for k, w in enumerate(net.get_weights()):
            out['w_td%d_%d' % (time_dim, k)] = w
        dlwp = models.DLWPNeuralNet(is_convolutional=True, is_recurrent=False, time_dim=time_dim, scaler_type=None,
                                    scale_targets=False)
        dlwp.model = _FakeKerasModel(net)
        for steps in (1, 5):
            for ss in (False, True):
                for ktd in (False, True):
                    key = 'y_td%d_s%d_ss%d_k%d' % (time_dim, steps, ss, ktd)
                    out[key] = dlwp.predict_timeseries(x0, steps, step_sequence=ss, keep_time_dim=ktd)
                    cases.append(key)
    out['cases'] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, 'rollout_neuralnet.npz'), **out)


class _SmallUnrolled(object):
    """A shared-weight net unrolled n times, like examples/train_functional.py:278-281, on top of an OSequential."""

    def __init__(self, net, n):
        self.net, self.n_outputs = net, n

    def forward(self, x):
        outs = [self.net.forward(x)]
        for _ in range(1, self.n_outputs):
            outs.append(self.net.forward(outs[-1]))
        return outs[0] if self.n_outputs == 1 else outs


def gen_rollout_functional(models):
    out = {}
    rng = np.random.RandomState(13)
    cases = []
    for time_dim in (1, 2):
        net = _small_seq(time_dim, 2, 6, 8, seed=30 + time_dim)
        x0 = rng.standard_normal((3, time_dim * 2, 6, 8)).astype(np.float32)
        out['x0_td%d' % time_dim] = x0
        for k, w in enumerate(net.get_weights()):
            out['w_td%d_%d' % (time_dim, k)] = w
        for n_steps in (1, 3):
            dlwp = models.DLWPFunctional(is_convolutional=True, is_recurrent=False, time_dim=time_dim)
            dlwp.model = _FakeKerasModel(_SmallUnrolled(net, n_steps))
            dlwp._n_steps = n_steps
            for steps in (1, 4, 7):
                for ktd in (False, True):
                    key = 'y_td%d_n%d_s%d_k%d' % (time_dim, n_steps, steps, ktd)
                    out[key] = dlwp.predict_timeseries(x0, steps, keep_time_dim=ktd)
                    cases.append(key)
    out['cases'] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, 'rollout_functional.npz'), **out)


def gen_torchnn(models_torch):
    """The reference's torch twin running "Net A" end to end (padding + conv + tanh + feedback loop) on CPU."""
    import torch
    models_torch.device = torch.device('cpu')
    torch.set_num_threads(1)  # deterministic summation order
    out = {}
    for tag, (C, H, W), n, steps, sub in (('small', (6, 23, 36), 2, 10, 1), ('full', (6, 91, 180), 1, 10, 6)):
        rng = np.random.RandomState(14)
        k1 = OO.glorot_uniform(rng, 3, 3, C, 32)
        b1 = (0.05 * rng.standard_normal(32)).astype(np.float32)
        k2 = OO.glorot_uniform(rng, 5, 5, 32, C)
        b2 = (0.05 * rng.standard_normal(C)).astype(np.float32)
        x0 = rng.standard_normal((n, C, H, W)).astype(np.float32)
        dlwp = models_torch.DLWPTorchNN(is_convolutional=True, is_recurrent=False, time_dim=1, scaler_type=None,
                                        scale_targets=False)
        layers = (
            ('CircularPad2d', ((2, 2, 0, 0),), None),
            ('ZeroPad2d', ((0, 0, 2, 2),), None),
            ('Conv2d', (C, 32, 3), {'dilation': 2, 'activation': 'tanh'}),
            ('CircularPad2d', ((2, 2, 0, 0),), None),
            ('ZeroPad2d', ((0, 0, 2, 2),), None),
            ('Conv2d', (32, C, 5), None),
        )
        dlwp.build_model(layers, 'Adam', 'MSELoss')
        with torch.no_grad():
            dlwp.layers[2].weight.copy_(torch.from_numpy(np.transpose(k1, (3, 2, 0, 1)).copy()))
            dlwp.layers[2].bias.copy_(torch.from_numpy(b1))
            dlwp.layers[5].weight.copy_(torch.from_numpy(np.transpose(k2, (3, 2, 0, 1)).copy()))
            dlwp.layers[5].bias.copy_(torch.from_numpy(b2))
        y = dlwp.predict_timeseries(x0, steps)
        out.update({'%s_x0' % tag: x0, '%s_k1' % tag: k1, '%s_b1' % tag: b1, '%s_k2' % tag: k2, '%s_b2' % tag: b2,
                    '%s_y' % tag: y[:, :, :, ::sub, ::sub], '%s_sub' % tag: np.int64(sub),
                    '%s_steps' % tag: np.int64(steps)})
    np.savez_compressed(os.path.join(HERE, 'torchnn_net_a.npz'), **out)


def main():
    install_stubs()
    util = load_ref('DLWP.util', 'DLWP/util.py')
    custom = load_ref('DLWP.custom', 'DLWP/custom.py')
    models = load_ref('DLWP.model.models', 'DLWP/model/models.py')
    models_torch = load_ref('DLWP.model.models_torch', 'DLWP/model/models_torch.py')
    only = set(sys.argv[1:])     # e.g. `make_golden.py padding3d recurrent`: regenerate just those fixtures
    gens = [('padding2d', lambda: gen_periodic_padding(custom)), ('row_conv', lambda: gen_row_conv(custom)),
            ('neuralnet', lambda: gen_rollout_neuralnet(models)), ('functional', lambda: gen_rollout_functional(models)),
            ('torchnn', lambda: gen_torchnn(models_torch)), ('padding3d', lambda: gen_padding_3d_and_fill(custom)),
            ('recurrent', lambda: gen_rollout_recurrent(models)), ('insolation', gen_insolation),
            ('acc_loss', lambda: gen_acc_loss(custom)), ('lat_loss', lambda: gen_lat_loss(custom)), ('estimator', lambda: gen_estimator(models, util)),
            ('series_generator', gen_series_generator), ('wrapper_pickles', lambda: gen_wrapper_pickles(models))]
    for name, fn in gens:
        if not only or name in only:
            fn()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print('%-28s %8d bytes' % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == '__main__':
    main()
