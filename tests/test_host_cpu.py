"""
CPU-only tests of the boundary: the C-ABI library loads and exports every declared symbol, the front end builds the
example nets unchanged and infers Keras shapes, the graph lowers to the expected fused ops, and argument errors match the
reference's exception types.  No kernel is launched here.
"""

import ctypes
import os
import pickle
import re

import numpy as np
import pytest

from oracle import layers as OL

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def nat():
    from dlwp_b200 import build
    build.build()  # no-op when up to date; nvcc cross-compiles without a GPU
    from dlwp_b200 import _native
    return _native


def test_library_exports_every_declared_symbol(nat):
    header = open(os.path.join(REPO, 'include', 'dlwp_b200.h')).read()
    declared = set(re.findall(r'\b(dlwp_[a-z0-9_]+)\s*\(', header))
    declared -= {'dlwp_stream_t'}
    assert declared == set(nat.SYMBOLS), declared ^ set(nat.SYMBOLS)
    lib = nat.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dlwp_abi_version() == nat.ABI_VERSION == 3
    assert lib.dlwp_kernel_launch_count() == 0


def test_struct_layouts_match_header(nat):
    assert ctypes.sizeof(nat.ConvDesc) == 22 * 4 + 6 * 8
    assert ctypes.sizeof(nat.BufferDesc) == 6 * 4
    assert ctypes.sizeof(nat.OpDesc) == 27 * 4
    assert ctypes.sizeof(nat.NetDesc) == 4 * 4 + 2 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(nat.PlanOptions) == 16 * 4
    assert nat.PlanOptions().tc_taps_in_k == -1 and nat.PlanOptions(math=1).math == 1


def test_library_has_tma_and_no_torch_dependency(nat):
    import subprocess
    out = subprocess.run(['ldd', nat.LIB_PATH], capture_output=True, text=True).stdout
    assert 'torch' not in out and 'libcuda.so' not in out


def test_build_model_validation_errors():
    from dlwp_b200.model import DLWPFunctional, DLWPNeuralNet
    m = DLWPNeuralNet(scaler_type=None, scale_targets=False)
    with pytest.raises(TypeError):
        m.build_model(layers='Conv2D')
    with pytest.raises(TypeError):
        m.build_model(layers=(('Conv2D', (1, 1), {}),), gpus=1.0)
    with pytest.raises(TypeError):
        m.build_model(layers=('Conv2D',))
    with pytest.raises(ValueError):
        m.build_model(layers=(('Conv2D', (1, 1)),))
    with pytest.raises(TypeError):
        m.build_model(layers=(('Conv2D', [1, 1], {}),))
    with pytest.raises(TypeError):
        m.build_model(layers=(('Conv2D', (1, 1), []),))
    with pytest.raises(AttributeError):
        m.build_model(layers=(('NoSuchLayer', None, None),))
    with pytest.raises(ValueError):
        DLWPNeuralNet(time_dim=0)
    with pytest.raises(ValueError):
        DLWPFunctional(time_dim=0)
    with pytest.raises(TypeError):
        DLWPFunctional().build_model(None, gpus='1')


def test_net_a_builds_from_reference_layer_tuples_and_lowers_to_two_fused_convs(nat):
    from dlwp_b200.engine import Lowering
    from tests.helpers import build_product_sequential
    dlwp = build_product_sequential(OL.net_a_layers())
    model = dlwp.model
    assert [l.__class__.__name__ for l in model.layers] == ['PeriodicPadding2D', 'ZeroPadding2D', 'Conv2D'] * 2
    assert model.layers[0].output_shape == (None, 6, 91, 184)
    assert model.layers[1].output_shape == (None, 6, 95, 184)
    assert model.layers[2].output_shape == (None, 32, 91, 180)
    assert model.output_shape == (None, 6, 91, 180)
    assert model.count_params() == 6566
    assert [w.shape for w in model.get_weights()] == [(3, 3, 6, 32), (32,), (5, 5, 32, 6), (6,)]
    low = Lowering(model)
    assert [o['kind'] for o in low.ops] == [nat.OP_CONV, nat.OP_CONV]       # no pad op, no copy
    for o in low.ops:
        assert (o['pad_t'], o['pad_b'], o['pad_l'], o['pad_r']) == (2, 2, 2, 2)
        assert (o['pad_mode_h'], o['pad_mode_w']) == (nat.PAD_ZERO, nat.PAD_PERIODIC)
    assert [b['kind'] for b in low.buffers] == [nat.BUF_INPUT, nat.BUF_INTERNAL, nat.BUF_OUTPUT]


def test_keras_shapes_follow_oracle_for_example_stack():
    """The full convolutional stack of examples/train.py:159-221 (non-recurrent part)."""
    from tests.helpers import build_product_sequential
    cf = 'channels_first'
    cs = (8, 36, 72)
    conv = lambda f, k, d, a: ('Conv2D', (f, k), {'dilation_rate': d, 'padding': 'valid', 'activation': a,
                                                  'data_format': cf})
    pp = lambda p: ('PeriodicPadding2D', ((0, p),), {'data_format': cf})
    zp = lambda p: ('ZeroPadding2D', ((p, 0),), {'data_format': cf})
    layers = (('PeriodicPadding2D', ((0, 2),), {'data_format': cf, 'input_shape': cs}), zp(2), conv(32, 3, 2, 'tanh'),
              ('MaxPooling2D', (2,), {'data_format': cf}), pp(1), zp(1), conv(64, 3, 1, 'tanh'),
              ('MaxPooling2D', (2,), {'data_format': cf}), pp(1), zp(1), conv(128, 3, 1, 'tanh'),
              ('UpSampling2D', (2,), {'data_format': cf}), pp(1), zp(1), conv(64, 3, 1, 'tanh'),
              ('UpSampling2D', (2,), {'data_format': cf}), pp(2), zp(2), conv(32, 3, 2, 'tanh'),
              pp(2), zp(2), conv(8, 5, 1, 'linear'), ('Reshape', (cs,), None))
    dlwp = build_product_sequential(layers)
    onet = OL.OSequential(layers)
    s = cs
    for kl, ol in zip(dlwp.model.layers, onet.layers):
        s = ol.output_shape(s)
        assert kl.output_shape == (None,) + tuple(s), kl.name
    assert dlwp.model.output_shape == (None,) + cs


def test_skip_model_lowering_has_no_pad_ops_and_elides_producer_copies(nat):
    from dlwp_b200.engine import Lowering
    from tests.helpers import build_functional_pair
    dlwp, _ = build_functional_pair((12, 16, 24), skip=True, integration_steps=2)
    assert dlwp._n_steps == 2
    low = Lowering(dlwp.model)
    kinds = [o['kind'] for o in low.ops]
    assert nat.OP_PAD not in kinds
    assert kinds.count(nat.OP_CONV) == 12 and kinds.count(nat.OP_COPY) == 4   # only the slice->concat skips copy
    assert sorted(o['weight_id'] for o in low.ops if o['kind'] == nat.OP_CONV) == sorted(list(range(6)) * 2)
    outs = [b for b in low.buffers if b['kind'] == nat.BUF_OUTPUT]
    assert sorted(b['output_index'] for b in outs) == [0, 1]


def test_padding_argument_normalisation_and_errors():
    from dlwp_b200.custom import PeriodicPadding2D, slice_layer
    assert PeriodicPadding2D(2).padding == ((2, 2), (2, 2))
    assert PeriodicPadding2D((0, 2)).padding == ((0, 0), (2, 2))
    assert PeriodicPadding2D(((1, 2), (3, 0))).padding == ((1, 2), (3, 0))
    with pytest.raises(ValueError):
        PeriodicPadding2D((1, 2, 3))
    with pytest.raises(ValueError):
        slice_layer(0, 4, axis=-1)
    with pytest.raises(ValueError):
        PeriodicPadding2D((1, 1), data_format='bogus')


def test_registry_lookup_order_matches_reference():
    from dlwp_b200 import util
    assert util.get_from_class('keras.layers', 'Conv2D').__name__ == 'Conv2D'
    with pytest.raises(AttributeError):
        util.get_from_class('keras.layers', 'PeriodicPadding2D')     # not a keras layer ...
    assert util.get_from_class('DLWP.custom', 'PeriodicPadding2D').pad_mode == 'periodic'  # ... but a DLWP.custom one
    names = util.get_classes('DLWP.custom')
    for n in ('PeriodicPadding2D', 'PeriodicPadding3D', 'FillPadding2D', 'TFPadding2D', 'RowConnected2D',
              'EarlyStoppingMin', 'RNNResetStates'):
        assert n in names
    for n in ('slice_layer', 'latitude_weighted_loss', 'anomaly_correlation_loss', 'lat_loss', 'acc_loss'):
        assert n in util.get_methods('DLWP.custom')


def test_wrapper_pickles_without_model_and_model_pickles_without_plan(tmp_path):
    from dlwp_b200 import util
    from tests.helpers import build_product_sequential
    dlwp = build_product_sequential(OL.net_a_layers((6, 12, 16)))
    w = dlwp.model.get_weights()
    base = str(tmp_path / 'm')
    util.save_model(dlwp, base)
    assert os.path.exists(base + '.keras') and os.path.exists(base + '.pkl')
    with open(base + '.pkl', 'rb') as f:
        bare = pickle.load(f)
    assert bare.model is None and bare.base_model is None
    back = util.load_model(base)
    assert back.time_dim == 1 and back.model is back.base_model
    for a, b in zip(back.model.get_weights(), w):
        np.testing.assert_array_equal(a, b)


def test_compat_aliases_install():
    import subprocess
    import sys
    code = ("import dlwp_b200.compat as c; c.install();"
            "from DLWP.model import DLWPNeuralNet, DLWPFunctional;"
            "from DLWP.custom import PeriodicPadding2D, RowConnected2D, slice_layer, RNNResetStates, EarlyStoppingMin,"
            " latitude_weighted_loss, anomaly_correlation_loss;"
            "from keras.layers import Input, ZeroPadding2D, ZeroPadding3D, Conv2D, ConvLSTM2D, MaxPooling2D,"
            " UpSampling2D, Reshape, concatenate;"
            "from keras.models import Model; from keras.callbacks import History, TensorBoard;"
            "from keras.regularizers import l2; from keras.losses import mean_squared_error;"
            "from DLWP.util import save_model, load_model, train_test_split_ind; print('ok')")
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=REPO)
    assert out.returncode == 0 and 'ok' in out.stdout, out.stderr


def test_latitude_weights_follow_reference_formula():
    from dlwp_b200.custom import latitude_weighted_loss
    from dlwp_b200.keras.losses import mean_squared_error
    lats = np.linspace(-90, 90, 7)
    f = latitude_weighted_loss(mean_squared_error, lats, (3, 7, 10), axis=-2, weighting='midlatitude')
    ref = np.cos(lats * np.pi / 180) + 0.5 * np.sin(lats * 2 * np.pi / 180) ** 2     # custom.py:976-978
    assert f.weights.shape == (7, 10)
    np.testing.assert_allclose(f.weights[:, 3], ref, rtol=1e-6, atol=1e-7)


def test_model_pickles_after_training_state_was_attached_and_with_latitude_loss(tmp_path):
    """Model.__getstate__ must drop every runtime handle (the training engine holds ctypes plan pointers)
    and latitude_weighted_loss / anomaly_correlation_loss must be picklable, or util.save_model fails after fit()."""
    import pickle
    from dlwp_b200 import util
    from dlwp_b200.custom import anomaly_correlation_loss, latitude_weighted_loss
    from dlwp_b200.keras.losses import mean_squared_error
    from dlwp_b200.model import DLWPNeuralNet
    layers = OL.net_a_layers((6, 10, 16))
    lats = np.linspace(-80, 80, 10)
    loss = latitude_weighted_loss(mean_squared_error, lats, (6, 10, 16), axis=-2, weighting='midlatitude')
    dlwp = DLWPNeuralNet(is_convolutional=True, time_dim=1, scaler_type=None, scale_targets=False)
    dlwp.build_model(layers, loss=loss, optimizer='adam')

    class FakeEngine(object):                   # what training._engine attaches: ctypes pointers cannot be pickled
        def __init__(self):
            self.plan = ctypes.pointer(ctypes.c_int(3))

        def close(self):
            pass
    dlwp.model._train_engine = FakeEngine()
    dlwp.model.history = FakeEngine()
    w0 = dlwp.model.get_weights()
    util.save_model(dlwp, str(tmp_path / 'm'))
    back = util.load_model(str(tmp_path / 'm'))
    for a, b in zip(w0, back.model.get_weights()):
        np.testing.assert_array_equal(a, b)
    l2 = back.model.loss
    y, yh = np.ones((2, 6, 10, 16), np.float32), np.zeros((2, 6, 10, 16), np.float32)
    assert abs(float(np.mean(l2(y, yh))) - float(np.mean(loss(y, yh)))) < 1e-7 and l2.__name__ == 'lat_loss'
    acc = pickle.loads(pickle.dumps(anomaly_correlation_loss(regularize_mean='mse')))
    rng = np.random.RandomState(0)
    a, b = rng.standard_normal((2, 3, 4, 5)), rng.standard_normal((2, 3, 4, 5))
    assert np.isfinite(acc(a, b)).all() and acc.__name__ == 'acc_loss'


def test_reference_pickled_wrappers_load_as_product_classes():
    """`<name>.pkl` of DLWP.util.save_model (util.py:143-149) as written by the REFERENCE's classes
    (tests/golden/make_golden.py:gen_wrapper_pickles): under compat.install() it unpickles into the product's classes with
    every attribute (fitted sklearn scalers included), and every submodule alias is the SAME module object as its
    dlwp_b200 original (no duplicate classes)."""
    import os
    import pickle
    import sys
    import dlwp_b200.compat
    import dlwp_b200.keras.engine
    import dlwp_b200.model
    dlwp_b200.compat.install()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'wrapper_pickles.npz'))
    nn = pickle.loads(bytes(g['neuralnet']))
    assert type(nn) is dlwp_b200.model.DLWPNeuralNet
    assert (nn.time_dim, nn.scaler_type, nn.scale_targets, nn.apply_same_y_scaling, nn.is_convolutional) == \
        (2, 'StandardScaler', True, False, False)
    assert nn.model is None and nn.base_model is None and nn._is_init_fit
    Xs, ys = nn.scaler_transform(g['X'], g['y'])
    np.testing.assert_allclose(Xs, g['Xs'], rtol=1e-12)
    np.testing.assert_allclose(ys, g['ys'], rtol=1e-12)
    fun = pickle.loads(bytes(g['functional']))
    assert type(fun) is dlwp_b200.model.DLWPFunctional and (fun.time_dim, fun._n_steps, fun.gpus) == (3, 4, 2)
    for alias, mod in (('DLWP.model.extensions', 'dlwp_b200.model.extensions'),
                       ('DLWP.model.models_torch', 'dlwp_b200.model.models_torch'),
                       ('DLWP.model.generators', 'dlwp_b200.model.generators'), ('keras.engine', 'dlwp_b200.keras.engine'),
                       ('keras.saving', 'dlwp_b200.keras.saving')):
        assert sys.modules[alias] is sys.modules[mod], alias
    from DLWP.model.extensions import TimeSeriesEstimator
    assert TimeSeriesEstimator is dlwp_b200.model.TimeSeriesEstimator


def test_training_loss_recognition():
    """Which compiled losses the device training path takes (training._check_compiled): 'mse', a latitude-weighted MSE, one
    anomaly-correlation loss for all outputs (custom.py:1036-1093); anything else is refused before any GPU work."""
    from dlwp_b200 import training
    from dlwp_b200.custom import acc_loss, anomaly_correlation_loss, latitude_weighted_loss
    from dlwp_b200.keras.losses import mean_absolute_error, mean_squared_error

    class M(object):
        _compile_kwargs = {}

    def ok(loss):
        m = M()
        m.loss = loss
        training._check_compiled(m)

    for good in ('mse', mean_squared_error, latitude_weighted_loss(mean_squared_error, np.linspace(-90, 90, 7), (2, 7, 9)),
                 acc_loss, 'acc_loss', anomaly_correlation_loss(regularize_mean=None), [acc_loss, acc_loss]):
        ok(good)
    assert training._loss_is_acc(anomaly_correlation_loss(mean=np.zeros((1, 2, 3, 4)), regularize_mean='mae'))
    assert not training._loss_is_acc('mse') and not training._loss_is_mse(acc_loss)
    for bad in ('mae', mean_absolute_error, [acc_loss, 'mse'], [acc_loss, anomaly_correlation_loss(regularize_mean='mae')]):
        with pytest.raises(NotImplementedError):
            ok(bad)


def test_recurrent_front_block_builds_and_lowers(nat):
    """examples/train.py:144-157: PeriodicPadding3D + ZeroPadding3D + ConvLSTM2D + Reshape in front of the conv stack.
    Per time step: input conv (pads fused), recurrent conv ('same', zero padded; absent at t = 0), one gate op."""
    from dlwp_b200.engine import Lowering
    from dlwp_b200.model import DLWPNeuralNet
    from tests.test_oracle_golden import small_recurrent_layers
    dlwp = DLWPNeuralNet(is_convolutional=True, is_recurrent=True, time_dim=3, scaler_type=None, scale_targets=False)
    dlwp.build_model(small_recurrent_layers(3), loss='mse', optimizer='adam')
    m = dlwp.model
    assert [l.output_shape for l in m.layers[:4]] == [(None, 3, 2, 6, 12), (None, 3, 2, 10, 12), (None, 3, 4, 6, 8),
                                                      (None, 12, 6, 8)]
    lstm = m.layers[2]
    assert [w.shape for w in lstm.get_weights()] == [(3, 3, 2, 16), (3, 3, 4, 16), (16,)]
    np.testing.assert_array_equal(lstm.get_weights()[2], np.repeat([0., 1., 0., 0.], 4))   # unit_forget_bias
    low = Lowering(m)
    kinds = [o['kind'] for o in low.ops]
    assert kinds == [nat.OP_CONV, nat.OP_LSTM, nat.OP_CONV, nat.OP_CONV, nat.OP_LSTM, nat.OP_CONV, nat.OP_CONV,
                     nat.OP_LSTM, nat.OP_CONV]
    first, rec = low.ops[0], low.ops[3]
    assert (first['pad_mode_w'], first['pad_mode_h'], first['pad_l'], first['pad_t'], first['dil_h']) == (
        nat.PAD_PERIODIC, nat.PAD_ZERO, 2, 2, 2)
    assert (rec['pad_mode_w'], rec['pad_l'], rec['pad_t'], rec['dil_h'], rec['dst_c0']) == (nat.PAD_ZERO, 1, 1, 1, 16)
    assert [o['src_c'] for o in low.ops if o['kind'] == nat.OP_LSTM] == [16, 32, 32]
    assert low.out_vals[0].shape == (3, 2, 6, 8)
    assert low.buffers[low.ops[1]['aux']]['C'] == 4        # the cell state survives buffer compaction
    import pickle
    m2 = pickle.loads(pickle.dumps(m))
    assert [w.shape for w in m2.get_weights()] == [w.shape for w in m.get_weights()]


def test_fill_and_tf_padding_lower_to_pad_ops(nat):
    from dlwp_b200.engine import Lowering
    from dlwp_b200.model import DLWPNeuralNet
    cf = 'channels_first'
    layers = (('FillPadding2D', ((1, 2),), {'data_format': cf, 'input_shape': (3, 9, 12)}),
              ('Conv2D', (5, 3), {'activation': 'tanh', 'data_format': cf}),
              ('TFPadding2D', ((1, 1),), {'data_format': cf, 'mode': 'SYMMETRIC'}),
              ('Conv2D', (3, 3), {'activation': 'linear', 'data_format': cf}))
    dlwp = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    dlwp.build_model(layers, loss='mse', optimizer='adam')
    low = Lowering(dlwp.model)
    assert [(o['kind'], o['pad_mode_h']) for o in low.ops] == [(nat.OP_PAD, nat.PAD_EDGE), (nat.OP_CONV, 0),
                                                               (nat.OP_PAD, nat.PAD_SYMMETRIC), (nat.OP_CONV, 0)]
    bad = (('TFPadding2D', ((1, 1),), {'data_format': cf, 'mode': 'CONSTANT', 'constant_values': 2.0,
                                       'input_shape': (3, 9, 12)}),
           ('Conv2D', (3, 3), {'data_format': cf}))
    dlwp = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    dlwp.build_model(bad, loss='mse', optimizer='adam')
    with pytest.raises(NotImplementedError):
        Lowering(dlwp.model)


def test_channels_last_model_lowers_with_layout_ops(nat):
    from dlwp_b200.engine import Lowering
    from dlwp_b200.model import DLWPNeuralNet
    from tests.test_channels_last_gpu import _layers
    dlwp = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    dlwp.build_model(_layers((12, 20, 5)), loss='mse', optimizer='adam')
    assert dlwp.model.output_shape == (None, 12, 20, 5)
    low = Lowering(dlwp.model)
    kinds = [o['kind'] for o in low.ops]
    assert kinds[0] == nat.OP_TO_NCHW and kinds[-1] == nat.OP_TO_NHWC and kinds.count(nat.OP_CONV) == 2
    assert low.out_vals[0].shape == (12, 20, 5)
    mixed = (('PeriodicPadding2D', ((0, 1),), {'data_format': 'channels_last', 'input_shape': (8, 8, 3)}),
             ('Conv2D', (4, 3), {'data_format': 'channels_first'}))
    dlwp = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    dlwp.build_model(mixed, loss='mse', optimizer='adam')
    with pytest.raises(NotImplementedError):
        Lowering(dlwp.model)


def test_time_series_estimator_host_loop_and_coordinates():
    """TimeSeriesEstimator (extensions.py:136-303) on the host loop, the network replaced by the oracle: values equal the
    numpy restatement (oracle/estimator.py), dimension names and coordinate values follow extensions.py:259-292."""
    from dlwp_b200.model import ArraySeriesGenerator, DLWPNeuralNet, TimeSeriesEstimator
    from oracle import estimator as OE
    rng = np.random.RandomState(0)
    varlev = ['z/500', 't/850', 'u/300']
    nt, Hh, Ww = 12, 6, 8
    data = rng.standard_normal((nt, 3, Hh, Ww)).astype(np.float32)
    times = np.datetime64('2003-03-01T00:00') + np.arange(nt) * np.timedelta64(6, 'h')
    lat, lon = np.linspace(80, -80, Hh), np.arange(0, 360, 45.)
    gen = ArraySeriesGenerator(data, times, lat, lon, varlev, ['z/500', 't/850', 'u/300'], ['z/500', 't/850'], 2, 1, 1, True)
    assert gen.convolution_shape == (8, Hh, Ww) and gen.output_convolution_shape == (2, Hh, Ww) and gen._n_sample == 10
    p, t = gen.generate([])
    assert p.shape == (10, 8, Hh, Ww) and t.shape == (10, 2, Hh, Ww)
    np.testing.assert_array_equal(t[0].reshape(1, 2, Hh, Ww)[0], data[2, :2])         # target = the step after the inputs
    w = rng.standard_normal((8, 2)).astype(np.float32)
    fn = lambda x: np.einsum('nchw,co->nohw', np.asarray(x, np.float32), w)
    dlwp = DLWPNeuralNet(is_convolutional=True, time_dim=2, scaler_type=None, scale_targets=False)
    dlwp.predict = lambda x, **kw: fn(x)
    est = TimeSeriesEstimator(dlwp, gen)
    est._device_ok = lambda: False
    out = est.predict(4)
    res, es, keep = OE.estimator_predict(fn, p, 4, 2, 1, ['z/500', 't/850', 'u/300', 'SOL'], ['z/500', 't/850'],
                                         gen.sample_times, np.timedelta64(6, 'h'), lat, lon, 1, True, False, True)
    ref = OE.estimator_series(res, 4, es, keep)
    np.testing.assert_array_equal(np.asarray(out.values), ref)
    assert np.isnan(ref[1, -1]).all() and np.isfinite(ref[1, :-1]).all()      # one sample per step runs out of data
    coords = out.coords if isinstance(out.coords, dict) else {k: v.values for k, v in out.coords.items()}
    assert tuple(out.dims) == ('f_hour', 'time', 'varlev', 'lat', 'lon')
    assert list(coords['f_hour']) == [np.timedelta64(6 * k, 'h') for k in (1, 2, 3, 4)]
    assert coords['time'][0] == times[1]                                       # the last input time of sample 0
    assert list(coords['varlev']) == ['z/500', 't/850']


def test_host_cpu_binding_is_a_no_op_without_a_gpu():
    """parallel.bind_host_near_gpu is an optimisation for one-process-per-GPU runs: without CUDA / NVML it changes nothing."""
    import os
    from dlwp_b200.parallel import bind_host_near_gpu
    before = os.sched_getaffinity(0)
    assert bind_host_near_gpu(0) == 0
    assert os.sched_getaffinity(0) == before
