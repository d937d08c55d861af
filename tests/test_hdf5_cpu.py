"""
Model / weight interchange (DLWP/util.py:126-192): the HDF5 container (dlwp_b200/hdf5.py) against a
file written by libhdf5 itself, writer -> reader round trips, and the Keras 2.2 model-file layout (keras/saving.py).
CPU only.
"""

import importlib.util
import json
import os
import pickle

import numpy as np
import pytest

from dlwp_b200 import hdf5


def _scipy_hdf5_file():
    spec = importlib.util.find_spec('scipy')
    if spec is None:
        return None
    p = os.path.join(os.path.dirname(spec.origin), 'io', 'matlab', 'tests', 'data', 'testhdf5_7.4_GLNX86.mat')
    return p if os.path.exists(p) else None


def test_reader_parses_a_file_written_by_libhdf5():
    """scipy ships a MATLAB v7.3 file = HDF5 written by libhdf5 1.x: 512-byte user block, superblock v0, B-tree + local
    heap group, contiguous float64 dataset, fixed-length string attribute.  Known contents: linspace(0, 2*pi, 9)."""
    path = _scipy_hdf5_file()
    if path is None:
        pytest.skip('scipy test data not installed')
    f = hdf5.File(path)
    assert (f.O, f.L, f.base) == (8, 8, 512)
    assert f.keys() == ['testdouble']
    d = f['testdouble']
    assert d.shape == (9, 1) and d.dtype == np.dtype('<f8')
    assert d.attrs['MATLAB_class'] == b'double'
    np.testing.assert_allclose(d.read()[:, 0], np.linspace(0, 2 * np.pi, 9), rtol=0, atol=1e-15)


def test_writer_reader_round_trip():
    rng = np.random.RandomState(0)
    w = hdf5.FileWriter()
    w.attrs['title'] = b'round trip'
    w.attrs['text'] = 'unicode é'
    w.attrs['count'] = np.int64(7)
    w.attrs['scale'] = 0.5
    w.attrs['names'] = [b'a', b'bcd', b'ef']
    w.attrs['empty'] = np.zeros((0,), np.float64)
    arrays = {}
    for i in range(75):                      # > 32 members: several symbol-table nodes under one B-tree node
        a = rng.standard_normal((3, i % 4 + 1)).astype(np.float32)
        arrays['w%03d' % i] = a
        w.create_dataset('many/w%03d' % i, a)
    w.create_dataset('a/b/c/kernel:0', np.arange(24, dtype=np.float64).reshape(2, 3, 4)).attrs['unit'] = b'K'
    w.create_dataset('ints', np.arange(-3, 4, dtype=np.int32))
    w.create_dataset('scalar', np.float32(2.5))
    w.create_dataset('strings', np.array([b'x', b'yy'], 'S2'))
    w.create_group('a').attrs['depth'] = np.int32(1)
    f = hdf5.File(w.tobytes())
    assert f.attrs['title'] == b'round trip' and f.attrs['text'].decode('utf8') == 'unicode é'
    assert f.attrs['count'] == 7 and f.attrs['scale'] == 0.5 and f.attrs['empty'].shape == (0,)
    assert list(f.attrs['names']) == [b'a', b'bcd', b'ef']
    assert sorted(f.keys()) == ['a', 'ints', 'many', 'scalar', 'strings']
    assert sorted(f['many'].keys()) == sorted(arrays)
    for k, a in arrays.items():
        np.testing.assert_array_equal(f['many/' + k].read(), a)
    c = f['a/b/c/kernel:0']
    assert c.attrs['unit'] == b'K' and c.dtype == np.float64
    np.testing.assert_array_equal(c.read(), np.arange(24.).reshape(2, 3, 4))
    np.testing.assert_array_equal(f['ints'].read(), np.arange(-3, 4))
    assert f['scalar'].read() == np.float32(2.5) and f['scalar'].shape == ()
    assert list(f['strings'].read()) == [b'x', b'yy']
    assert f['a'].attrs['depth'] == 1
    with pytest.raises(KeyError):
        f['a/nope']


def test_writer_emits_the_structures_libhdf5_expects():
    """Byte-level checks of the writer against the format specification (independent of this package's reader)."""
    w = hdf5.FileWriter()
    w.create_dataset('x', np.arange(4, dtype=np.float32))
    b = w.tobytes()
    assert b[:8] == b'\x89HDF\r\n\x1a\n' and b[8] == 0            # superblock version 0
    assert (b[13], b[14]) == (8, 8)                                 # sizes of offsets / lengths
    import struct
    base, free, eof, drv = struct.unpack_from('<QQQQ', b, 24)
    assert base == 0 and free == hdf5.UNDEF and eof == len(b) and drv == hdf5.UNDEF
    name_off, ohdr, cache, _ = struct.unpack_from('<QQII', b, 56)
    bt, hp = struct.unpack_from('<QQ', b, 80)
    assert cache == 1 and b[bt:bt + 4] == b'TREE' and b[hp:hp + 4] == b'HEAP' and b[ohdr] == 1
    assert ohdr % 8 == 0 and bt % 8 == 0 and hp % 8 == 0
    snod, = struct.unpack_from('<Q', b, bt + 24 + 8)
    assert b[snod:snod + 4] == b'SNOD' and struct.unpack_from('<H', b, snod + 6)[0] == 1
    raw = np.arange(4, dtype=np.float32).tobytes()
    assert b.count(raw) == 1 and b.index(raw) % 8 == 0              # contiguous, aligned dataset storage


def _net_a(shape=(6, 16, 32)):
    from dlwp_b200.model import DLWPNeuralNet
    from oracle import layers as OL
    d = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    d.build_model(OL.net_a_layers(shape), loss='mse', optimizer='adam')
    return d


def test_save_model_writes_keras_hdf5_and_load_model_restores_it(tmp_path):
    from dlwp_b200 import util
    d = _net_a()
    base = str(tmp_path / 'm')
    util.save_model(d, base)
    f = hdf5.File(base + '.keras')
    assert f.attrs['keras_version'] == b'2.2.4' and f.attrs['backend'] == b'tensorflow'
    cfg = json.loads(f.attrs['model_config'].decode('utf8'))
    assert cfg['class_name'] == 'Sequential'
    recs = cfg['config']['layers']
    assert [r['class_name'] for r in recs] == ['PeriodicPadding2D', 'ZeroPadding2D', 'Conv2D'] * 2
    assert recs[0]['config']['batch_input_shape'] == [None, 6, 16, 32]
    assert recs[2]['config']['dilation_rate'] == [2, 2] and recs[2]['config']['activation'] == 'tanh'
    mw = f['model_weights']
    assert [n.decode() for n in mw.attrs['layer_names']] == [l.name for l in d.model.layers]
    k = mw['conv2d_1/conv2d_1/kernel:0'] if 'conv2d_1' in mw.keys() else None
    name = d.model.layers[2].name
    assert [n.decode() for n in mw[name].attrs['weight_names']] == [name + '/kernel:0', name + '/bias:0']
    np.testing.assert_array_equal(mw['%s/%s/kernel:0' % (name, name)].read(), d.model.layers[2].kernel)
    tc = json.loads(f.attrs['training_config'].decode('utf8'))
    assert tc['optimizer_config']['class_name'] == 'Adam' and tc['loss'] == 'mse'
    assert k is None or k.shape == (3, 3, 6, 32)
    d2 = util.load_model(base)
    assert [l.name for l in d2.model.layers] == [l.name for l in d.model.layers]
    for a, b in zip(d.model.get_weights(), d2.model.get_weights()):
        np.testing.assert_array_equal(a, b)
    assert d2.model.optimizer.lr == d.model.optimizer.lr and d2.model.loss == 'mse'
    assert d2.time_dim == d.time_dim and d2.model is d2.base_model


def test_functional_unet_with_shared_layers_round_trips(tmp_path):
    """skip_model of examples/train_functional.py:248-275 unrolled twice: shared layers (two inbound nodes each),
    slice_layer Lambdas, concatenate."""
    from dlwp_b200 import util
    from tests.helpers import build_functional_pair
    d, _ = build_functional_pair((4, 16, 32), skip=True, integration_steps=2)
    base = str(tmp_path / 'u')
    util.save_model(d, base)
    cfg = json.loads(hdf5.File(base + '.keras').attrs['model_config'].decode('utf8'))
    assert cfg['class_name'] == 'Model' and len(cfg['config']['output_layers']) == 2
    conv = [r for r in cfg['config']['layers'] if r['class_name'] == 'Conv2D'][0]
    assert len(conv['inbound_nodes']) == 2                               # one node per unrolled application
    lam = [r for r in cfg['config']['layers'] if r['class_name'] == 'Lambda']
    assert len(lam) == 4 and lam[0]['config']['function'][2] == [1, 16, 0, None]
    d2 = util.load_model(base)
    assert len(d2.model.outputs) == 2 and d2.model.output_shape == d.model.output_shape
    assert [l.__class__.__name__ for l in d2.model.layers] == [l.__class__.__name__ for l in d.model.layers]
    for a, b in zip(d.model.get_weights(), d2.model.get_weights()):
        np.testing.assert_array_equal(a, b)
    from dlwp_b200.engine import Lowering
    assert [o['kind'] for o in Lowering(d2.model).ops] == [o['kind'] for o in Lowering(d.model).ops]


def test_recurrent_and_row_connected_models_round_trip(tmp_path):
    from dlwp_b200.model import DLWPNeuralNet
    from tests.test_oracle_golden import small_recurrent_layers
    d = DLWPNeuralNet(is_convolutional=True, is_recurrent=True, time_dim=2, scaler_type=None, scale_targets=False)
    d.build_model(small_recurrent_layers(2), loss='mse', optimizer='adam')
    p = str(tmp_path / 'r.keras')
    d.model.save(p)
    f = hdf5.File(p)
    lstm = d.model.layers[2].name
    assert [n.decode() for n in f['model_weights'][lstm].attrs['weight_names']] == [
        lstm + '/kernel:0', lstm + '/recurrent_kernel:0', lstm + '/bias:0']
    from dlwp_b200.keras.models import load_model
    m2 = load_model(p)
    assert m2.output_shape == d.model.output_shape
    for a, b in zip(d.model.get_weights(), m2.get_weights()):
        np.testing.assert_array_equal(a, b)
    cf = 'channels_first'
    d = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    d.build_model((('RowConnected2D', (3, 3), {'data_format': cf, 'input_shape': (2, 7, 9)}),), loss='mse', optimizer='adam')
    p = str(tmp_path / 'rc.keras')
    d.model.save(p)
    m2 = load_model(p)
    assert m2.layers[0].__class__.__name__ == 'RowConnected2D' and m2.get_weights()[0].shape == (5, 3, 3, 2, 3)


def test_loads_a_file_laid_out_like_keras_2_2_writes_it(tmp_path):
    """A model file assembled the way Keras 2.2.4 + h5py lay it out for a reference-built net (full layer configs with
    initializer / regularizer dicts, dtype, the marshalled slice_layer Lambda) -- built with the HDF5 writer, not Keras."""
    glorot = {'class_name': 'VarianceScaling', 'config': {'scale': 1.0, 'mode': 'fan_avg', 'distribution': 'uniform',
                                                          'seed': None}}

    def conv(name, filters, k, act, first=False):
        c = {'name': name, 'trainable': True, 'filters': filters, 'kernel_size': [k, k], 'strides': [1, 1],
             'padding': 'valid', 'data_format': 'channels_first', 'dilation_rate': [1, 1], 'activation': act,
             'use_bias': True, 'kernel_initializer': glorot, 'bias_initializer': {'class_name': 'Zeros', 'config': {}},
             'kernel_regularizer': {'class_name': 'L1L2', 'config': {'l1': 0.0, 'l2': 9.999999747378752e-05}},
             'bias_regularizer': None, 'activity_regularizer': None, 'kernel_constraint': None, 'bias_constraint': None}
        return c
    pad = lambda name, p: {'name': name, 'trainable': True, 'padding': p, 'data_format': 'channels_first'}
    code = 'YwEAAAAAAAAAAwAAAAQAAAATAAAA...'     # marshalled byte code (never executed by the loader)
    layers = [
        {'name': 'input_0', 'class_name': 'InputLayer', 'inbound_nodes': [],
         'config': {'batch_input_shape': [None, 4, 8, 12], 'dtype': 'float32', 'sparse': False, 'name': 'input_0'}},
        {'name': 'periodic_padding2d_1', 'class_name': 'PeriodicPadding2D', 'config': pad('periodic_padding2d_1', [[0, 0], [1, 1]]),
         'inbound_nodes': [[['input_0', 0, 0, {}]]]},
        {'name': 'zero_padding2d_1', 'class_name': 'ZeroPadding2D', 'config': pad('zero_padding2d_1', [[1, 1], [0, 0]]),
         'inbound_nodes': [[['periodic_padding2d_1', 0, 0, {}]]]},
        {'name': 'conv2d_1', 'class_name': 'Conv2D', 'config': conv('conv2d_1', 8, 3, 'tanh'),
         'inbound_nodes': [[['zero_padding2d_1', 0, 0, {}]]]},
        {'name': 'lambda_1', 'class_name': 'Lambda', 'inbound_nodes': [[['conv2d_1', 0, 0, {}]]],
         'config': {'name': 'lambda_1', 'trainable': True, 'function': [code, None, [1, 4, 0, None]],
                    'function_type': 'lambda', 'output_shape': None, 'output_shape_type': 'raw', 'arguments': {}}},
        {'name': 'lambda_2', 'class_name': 'Lambda', 'inbound_nodes': [[['conv2d_1', 0, 0, {}]]],
         'config': {'name': 'lambda_2', 'trainable': True, 'function': [code, None, [1, 8, 4, None]],
                    'function_type': 'lambda', 'output_shape': None, 'output_shape_type': 'raw', 'arguments': {}}},
        {'name': 'concatenate_1', 'class_name': 'Concatenate', 'config': {'name': 'concatenate_1', 'trainable': True, 'axis': 1},
         'inbound_nodes': [[['lambda_2', 0, 0, {}], ['lambda_1', 0, 0, {}]]]},
    ]
    cfg = {'class_name': 'Model', 'config': {'name': 'model_1', 'layers': layers, 'input_layers': [['input_0', 0, 0]],
                                             'output_layers': [['concatenate_1', 0, 0]]}}
    rng = np.random.RandomState(2)
    kern, bias = rng.standard_normal((3, 3, 4, 8)).astype(np.float32), rng.standard_normal(8).astype(np.float32)
    w = hdf5.FileWriter()
    w.attrs['keras_version'] = b'2.2.4'
    w.attrs['backend'] = b'tensorflow'
    w.attrs['model_config'] = json.dumps(cfg).encode('utf8')
    w.attrs['training_config'] = json.dumps({
        'optimizer_config': {'class_name': 'Adam', 'config': {'lr': 0.0005, 'beta_1': 0.9, 'beta_2': 0.999, 'decay': 0.0,
                                                               'epsilon': 1e-07, 'amsgrad': False}},
        'loss': 'mean_squared_error', 'metrics': ['mae'], 'sample_weight_mode': None, 'loss_weights': None}).encode('utf8')
    g = w.create_group('model_weights')
    names = [l['name'] for l in layers]
    g.attrs['layer_names'] = [n.encode() for n in names]
    g.attrs['backend'] = b'tensorflow'
    g.attrs['keras_version'] = b'2.2.4'
    for n in names:
        lg = g.create_group(n)
        if n == 'conv2d_1':
            lg.attrs['weight_names'] = [b'conv2d_1/kernel:0', b'conv2d_1/bias:0']
            lg.create_dataset('conv2d_1/kernel:0', kern)
            lg.create_dataset('conv2d_1/bias:0', bias)
        else:
            lg.attrs['weight_names'] = np.zeros((0,), np.float64)
    path = str(tmp_path / 'ref.keras')
    w.save(path)
    from dlwp_b200.keras.models import load_model
    m = load_model(path)
    assert m.output_shape == (None, 8, 8, 12)
    np.testing.assert_array_equal(m.get_layer('conv2d_1').kernel, kern)
    np.testing.assert_array_equal(m.get_layer('conv2d_1').bias, bias)
    assert abs(m.get_layer('conv2d_1').kernel_regularizer.l2 - 1e-4) < 1e-9
    sl = m.get_layer('lambda_2')
    assert (sl.start, sl.end, sl.step, sl.axis) == (4, 8, None, 1)
    assert m.optimizer.lr == 0.0005 and m.loss == 'mean_squared_error'
    bad = json.loads(json.dumps(cfg))
    bad['config']['layers'][4]['config']['function'] = [code, None, None]      # a Lambda that is not a slice_layer
    w.attrs['model_config'] = json.dumps(bad).encode('utf8')
    w.save(path)
    with pytest.raises(NotImplementedError):
        load_model(path)


def test_legacy_pickle_container_still_loads(tmp_path):
    d = _net_a()
    p = str(tmp_path / 'old.keras')
    with open(p, 'wb') as f:
        pickle.dump({'format': 'dlwp_b200.keras.v1', 'model': d.model}, f, protocol=pickle.HIGHEST_PROTOCOL)
    from dlwp_b200.keras.models import load_model
    m = load_model(p)
    for a, b in zip(d.model.get_weights(), m.get_weights()):
        np.testing.assert_array_equal(a, b)


def test_save_and_load_weights_only(tmp_path):
    d, e = _net_a(), _net_a()
    p = str(tmp_path / 'w.h5')
    d.model.save_weights(p)
    e.model.load_weights(p)
    for a, b in zip(d.model.get_weights(), e.model.get_weights()):
        np.testing.assert_array_equal(a, b)
