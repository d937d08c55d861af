"""
`TimeSeriesEstimator.predict` (DLWP/model/extensions.py:136-303) with the predictor array resident on
the GPU, against the numpy restatement of the reference's loop (oracle/estimator.py -- unpinned: the reference needs
xarray) driven by the float64 oracle net.  Configurations: fewer output than input time steps with an input the model does
not predict and the insolation forcing; more output than input time steps (first / last times preferred); imputing.
"""

import numpy as np
import pytest

from oracle import estimator as OE
from oracle import layers as OL
from tests.helpers import build_product_sequential

pytestmark = pytest.mark.gpu
H, W = 12, 16


def _setup(t_in, t_out, in_vl, out_vl, sol, seed=0):
    from dlwp_b200.model import ArraySeriesGenerator
    rng = np.random.RandomState(seed)
    varlev = ['z/500', 't/850', 'u/300']
    nt = 14
    data = rng.standard_normal((nt, 3, H, W)).astype(np.float32)
    times = np.datetime64('2003-03-01T00:00') + np.arange(nt) * np.timedelta64(6, 'h')
    lat, lon = np.linspace(80, -80, H), np.arange(0, 360, 360. / W)
    gen = ArraySeriesGenerator(data, times, lat, lon, varlev, in_vl, out_vl, t_in, t_out, 1, sol)
    cf = 'channels_first'
    cin, cout = gen.convolution_shape[0], gen.output_convolution_shape[0]
    layers = (('PeriodicPadding2D', ((0, 1),), {'data_format': cf, 'input_shape': gen.convolution_shape}),
              ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
              ('Conv2D', (8, 3), {'activation': 'tanh', 'data_format': cf}),
              ('PeriodicPadding2D', ((0, 1),), {'data_format': cf}),
              ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
              ('Conv2D', (cout, 3), {'activation': 'linear', 'data_format': cf}))
    dlwp = build_product_sequential(layers, time_dim=t_in)
    net = OL.OSequential(layers)
    OL.init_weights(net.conv_layers, seed=seed + 1, bias_scale=0.05)
    dlwp.model.set_weights(net.get_weights())
    assert cin == t_in * (len(in_vl) + (1 if sol else 0))
    return dlwp, net, gen


@pytest.mark.parametrize('t_in,t_out,in_vl,out_vl,sol,impute,first', [
    (2, 1, ['z/500', 't/850', 'u/300'], ['z/500', 't/850'], True, False, True),
    (2, 2, ['z/500', 't/850'], ['z/500', 't/850'], False, False, True),
    (1, 2, ['z/500', 't/850'], ['t/850', 'z/500'], False, False, True),
    (1, 2, ['z/500', 't/850'], ['z/500', 't/850'], True, False, False),
    (2, 1, ['z/500', 'u/300'], ['z/500'], True, True, True),
])
def test_estimator_matches_restated_reference_loop(t_in, t_out, in_vl, out_vl, sol, impute, first):
    from dlwp_b200.model import TimeSeriesEstimator
    dlwp, net, gen = _setup(t_in, t_out, in_vl, out_vl, sol)
    est = TimeSeriesEstimator(dlwp, gen)
    steps = 5
    got = est.predict(steps, impute=impute, prefer_first_times=first)
    p, _ = gen.generate([], scale_and_impute=False)
    in_names = list(in_vl) + (['SOL'] if sol else [])
    fn = lambda x: net.forward(np.asarray(x, np.float64)).astype(np.float32)
    res, es, keep = OE.estimator_predict(fn, p, steps, t_in, t_out, in_names, out_vl, gen.sample_times,
                                         gen.times[1] - gen.times[0], gen.lat, gen.lon, 1, sol, impute, first)
    ref = OE.estimator_series(res, steps, es, keep, False, first)
    vals = np.asarray(got.values)
    assert tuple(got.dims) == ('f_hour', 'time', 'varlev', 'lat', 'lon')
    assert vals.shape == ref.shape == (steps, gen._n_sample, len(out_vl), H, W)
    np.testing.assert_array_equal(np.isnan(vals), np.isnan(ref))          # the same samples run out of data
    assert np.isfinite(ref[:, 0]).all()
    ok = np.isfinite(ref)
    assert np.abs(vals[ok] - ref[ok]).max() <= 3e-5 * np.abs(ref[ok]).max()
    kt = est.predict(steps, impute=impute, prefer_first_times=first, keep_time_dim=True)
    assert tuple(kt.dims) == ('f_hour', 'time', 'time_step', 'varlev', 'lat', 'lon')
    okk = np.isfinite(res)
    assert np.abs(np.asarray(kt.values)[okk] - res[okk]).max() <= 3e-5 * np.abs(res[okk]).max()
    # f_hour coordinates (extensions.py:283-285): multiples of the sample spacing
    fh = np.asarray(got.coords['f_hour']) if isinstance(got.coords, dict) else got.coords['f_hour'].values
    assert fh[0] == np.timedelta64(6, 'h') and len(fh) == steps


def test_estimator_argument_errors():
    from dlwp_b200.model import TimeSeriesEstimator
    dlwp, _, gen = _setup(1, 1, ['z/500'], ['z/500'], False)
    with pytest.raises(TypeError):
        TimeSeriesEstimator(object(), gen)
    with pytest.raises(TypeError):
        TimeSeriesEstimator(dlwp, object())
    with pytest.raises(ValueError):
        TimeSeriesEstimator(dlwp, gen).predict(0)


def test_device_series_generator_matches_host_assembly_and_trains():
    """SURVEY.md 8f rank 4 (second half): SeriesDataGenerator batch assembly on the GPU (dlwp_gather_series) -- bit identical
    to the host assembly of ArraySeriesGenerator.generate (generators.py:529-605 restated), insolation channel and target
    sequence included; fit_generator consumes its CUDA tensors directly."""
    import torch
    from dlwp_b200.model import ArraySeriesGenerator, DeviceSeriesGenerator
    rng = np.random.RandomState(5)
    varlev = ['z/500', 't/850', 'u/300']
    nt = 20
    data = rng.standard_normal((nt, 3, H, W)).astype(np.float32)
    times = np.datetime64('2003-03-01T00:00') + np.arange(nt) * np.timedelta64(6, 'h')
    lat, lon = np.linspace(80, -80, H), np.arange(0, 360, 360. / W)
    series = ArraySeriesGenerator(data, times, lat, lon, varlev, ['u/300', 'z/500'], ['z/500'], 2, 2, 1, True)
    gen = DeviceSeriesGenerator(series, batch_size=4, sequence=3, shuffle=False)
    assert gen._n_sample == nt - 2 - 2 * 3 - 1 + 2 and len(gen) == (gen._n_sample + 3) // 4
    p_all, _ = series.generate([])
    for b in (0, len(gen) - 1):
        X, ys = gen[b]
        assert X.is_cuda and len(ys) == 3
        idx = np.arange(b * 4, min((b + 1) * 4, gen._n_sample))
        np.testing.assert_array_equal(X.cpu().numpy(), p_all[idx])
        for s, y in enumerate(ys):
            ref = np.stack([data[idx + 2 + 2 * s + n][:, [0]] for n in range(2)], axis=1).reshape(len(idx), 2, H, W)
            np.testing.assert_array_equal(y.cpu().numpy(), ref)
    # fit_generator straight from device batches: a 1-output net on sequence=None
    gen1 = DeviceSeriesGenerator(ArraySeriesGenerator(data, times, lat, lon, varlev, ['z/500', 't/850'], ['z/500', 't/850'],
                                                      1, 1, 1, False), batch_size=6, shuffle=True)
    cf = 'channels_first'
    layers = (('PeriodicPadding2D', ((0, 1),), {'data_format': cf, 'input_shape': gen1.convolution_shape}),
              ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
              ('Conv2D', (2, 3), {'activation': 'linear', 'data_format': cf}))
    dlwp = build_product_sequential(layers)
    h = dlwp.model.fit_generator(gen1, epochs=3, verbose=0)
    loss = h.history['loss']
    assert len(loss) == 3 and np.isfinite(loss).all() and loss[-1] < loss[0]
    assert isinstance(gen1[0][0], torch.Tensor)


def test_device_series_generator_matches_reference_generator():
    """dlwp_gather_series / DeviceSeriesGenerator vs the reference's own SeriesDataGenerator.__getitem__ and generate
    (generators.py:529-640, executed by tests/golden/make_golden.py:gen_series_generator): batch count, the second batch
    (predictors with the insolation channel, every target of the sequence) and all samples."""
    import os
    from dlwp_b200.model import DeviceSeriesGenerator
    from tests.test_oracle_golden import series_generator_cases
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    for g, key, series, (t_in, t_out, seq, interval, sol, batch, n_sample, n_batches) in series_generator_cases(golden):
        gen = DeviceSeriesGenerator(series, batch_size=batch, sequence=seq or None, shuffle=False)
        assert gen._n_sample == n_sample and len(gen) == n_batches, key
        X, ys = gen[1]
        ys = ys if seq else [ys]
        np.testing.assert_allclose(X.cpu().numpy(), g[key + '/xb'], rtol=0, atol=1e-6, err_msg=key)
        for k, y in enumerate(ys):
            np.testing.assert_array_equal(y.cpu().numpy(), g[key + '/yb%d' % k], err_msg=key)
        full = DeviceSeriesGenerator(series, batch_size=n_sample, sequence=seq or None, shuffle=False)
        X, ys = full[0]
        ys = ys if seq else [ys]
        np.testing.assert_allclose(X.cpu().numpy(), g[key + '/p'], rtol=0, atol=1e-6, err_msg=key)
        for k, y in enumerate(ys):
            np.testing.assert_array_equal(y.cpu().numpy(), g[key + '/t%d' % k], err_msg=key)
