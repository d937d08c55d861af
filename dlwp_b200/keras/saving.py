"""
The reference's on-disk network format: what `keras.Model.save('<name>.keras')` writes and `keras.models.load_model`
reads (DLWP/util.py:139-141, :173), restated for Keras 2.2.x on top of dlwp_b200.hdf5 (h5py is not available offline).

File layout (keras/engine/saving.py of Keras 2.2; Keras is not vendored in /root/reference, so this layout is restated
from its published format -- interchange with real Keras files is UNPINNED, the HDF5 container itself is pinned against
a libhdf5-written file, see dlwp_b200/hdf5.py):

    /                      attrs  keras_version, backend, model_config (JSON), training_config (JSON, if compiled)
    /model_weights         attrs  layer_names (fixed-length strings), backend, keras_version
    /model_weights/<layer> attrs  weight_names, e.g. [b'conv2d_1/kernel:0', b'conv2d_1/bias:0']
    /model_weights/<layer>/<layer>/kernel:0   float32 dataset, Keras layout (kh, kw, Cin, Cout)

`model_config` is {"class_name": "Sequential" | "Model", "config": ...} with per-layer {"class_name", "config"} records;
functional models add "inbound_nodes" / "input_layers" / "output_layers".  Layer classes resolve in `keras.layers`, then in
`DLWP.custom` (the custom_objects sweep of DLWP/util.py:171-173), then in the caller's custom_objects.

`slice_layer` (DLWP/custom.py:675-692) is a Lambda around a closure; Keras serialises the closure's cell contents in
co_freevars order (axis, end, start, step) next to marshalled byte code.  The byte code is never executed here: a Lambda
whose closure has that 4-tuple shape is rebuilt as a ChannelSlice.  Files written here store ChannelSlice the same way
with an empty code string -- stock Keras cannot unmarshal that, so U-Nets with skip connections travel reference ->
dlwp_b200 but not back (Sequential nets travel both ways).
"""

import inspect
import json
import pickle
import warnings

import numpy as np

from .. import hdf5

KERAS_VERSION = b'2.2.4'
BACKEND = b'tensorflow'


# ---------------------------------------------------------------------------------------------------------------------
# config <-> layer
# ---------------------------------------------------------------------------------------------------------------------

def _jsonable(v):
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (tuple, list)):
        return [_jsonable(x) for x in v]
    if isinstance(v, dict):
        return {k: _jsonable(x) for k, x in v.items()}
    return v


def _regularizer_config(r):
    if r is None:
        return None
    return {'class_name': 'L1L2', 'config': {'l1': float(getattr(r, 'l1', 0.)), 'l2': float(getattr(r, 'l2', 0.))}}


def layer_config(layer):
    from .layers import ChannelSlice
    cfg = dict(layer.get_config())
    cfg.setdefault('name', layer.name)
    cfg.setdefault('trainable', getattr(layer, 'trainable', True))
    for k in ('kernel_regularizer', 'recurrent_regularizer', 'bias_regularizer'):
        if hasattr(layer, k):
            cfg[k] = _regularizer_config(getattr(layer, k))
    if isinstance(layer, ChannelSlice):
        # keras Lambda config; the closure slot carries (axis, end, start, step) like the reference's slice_func
        cfg.update({'function': ['', None, [layer.axis, layer.end, layer.start, layer.step]], 'function_type': 'lambda',
                    'output_shape': None, 'output_shape_type': 'raw', 'arguments': {}})
        return 'Lambda', _jsonable(cfg)
    return layer.__class__.__name__, _jsonable(cfg)


def _registry(custom_objects):
    from .. import custom
    from . import layers as KL
    reg = {}
    for mod in (custom, KL):      # keras.layers wins over DLWP.custom on a name clash, as in models.py:97-103
        for k in dir(mod):
            v = getattr(mod, k)
            if isinstance(v, type) and not k.startswith('_'):
                reg[k] = v
    reg.update(custom_objects or {})
    return reg


def _deserialize_value(k, v):
    from . import regularizers
    if k.endswith('_regularizer') and isinstance(v, dict):
        c = v.get('config', {})
        return regularizers.L1L2(l1=c.get('l1', 0.), l2=c.get('l2', 0.))
    if k.endswith('_initializer') and isinstance(v, dict):
        return str(v.get('class_name', 'glorot_uniform'))
    if k.endswith('_constraint'):
        return None
    if isinstance(v, list) and k in ('kernel_size', 'strides', 'dilation_rate', 'pool_size', 'size', 'target_shape',
                                     'batch_input_shape', 'padding'):
        return tuple(tuple(x) if isinstance(x, list) else x for x in v)
    return v


def layer_from_config(class_name, config, registry):
    from .layers import ChannelSlice, Lambda
    config = dict(config)
    if class_name == 'Lambda':
        fn = config.get('function')
        closure = fn[2] if isinstance(fn, (list, tuple)) and len(fn) == 3 else None
        if isinstance(closure, (list, tuple)) and len(closure) == 4 and isinstance(closure[0], int):
            axis, end, start, step = closure
            return ChannelSlice(start, end, step, axis, name=config.get('name'))
        raise NotImplementedError('Lambda layer %r: only the slice_layer closure of DLWP/custom.py:675-692 can be rebuilt '
                                  '(arbitrary marshalled byte code is never executed)' % config.get('name'))
    if class_name not in registry:
        raise ValueError('Unknown layer: ' + class_name)
    cls = registry[class_name]
    if cls is Lambda:
        raise NotImplementedError('Lambda layers cannot be rebuilt')
    accepted = set()
    for klass in inspect.getmro(cls):
        if klass is object:
            continue
        try:
            sig = inspect.signature(klass.__init__)
        except (TypeError, ValueError):
            continue
        accepted |= {n for n, p in sig.parameters.items() if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)}
    accepted.discard('self')
    kwargs = {k: _deserialize_value(k, v) for k, v in config.items() if k in accepted}
    if 'batch_input_shape' in kwargs and kwargs['batch_input_shape'] is not None:
        kwargs['batch_input_shape'] = tuple(kwargs['batch_input_shape'])
    return cls(**kwargs)


# ---------------------------------------------------------------------------------------------------------------------
# model_config
# ---------------------------------------------------------------------------------------------------------------------

def model_config(model):
    from .engine import InputLayer, Sequential
    if isinstance(model, Sequential):
        layers = []
        for i, l in enumerate(model.layers):
            cn, cfg = layer_config(l)
            if i == 0:
                cfg['batch_input_shape'] = _jsonable(model.inputs[0].shape)
                cfg.setdefault('dtype', 'float32')
            layers.append({'class_name': cn, 'config': cfg})
        return {'class_name': 'Sequential', 'config': {'name': model.name, 'layers': layers}}
    # functional: every call of a (possibly shared) layer is one inbound node
    node_index = {}
    for l in model.layers:
        for k, node in enumerate(l._inbound_nodes):
            node_index[id(node)] = k
    used = {id(n) for n in model._nodes}
    recs = []
    for l in model.layers:
        cn, cfg = layer_config(l)
        if isinstance(l, InputLayer):
            cn = 'InputLayer'
            cfg = {'batch_input_shape': _jsonable(l._batch_input_shape), 'dtype': 'float32', 'sparse': False, 'name': l.name}
        inbound = []
        for node in l._inbound_nodes:
            if id(node) not in used or not node.inputs:
                continue
            inbound.append([[t._node.layer.name, node_index[id(t._node)], 0, {}] for t in node.inputs])
        recs.append({'name': l.name, 'class_name': cn, 'config': cfg, 'inbound_nodes': inbound})

    def ref(t):
        return [t._node.layer.name, node_index[id(t._node)], 0]
    return {'class_name': 'Model', 'config': {'name': model.name, 'layers': recs,
                                              'input_layers': [ref(t) for t in model.inputs],
                                              'output_layers': [ref(t) for t in model.outputs]}}


def model_from_config(cfg, custom_objects=None):
    from .engine import Input, Model, Sequential
    reg = _registry(custom_objects)
    cls, c = cfg['class_name'], cfg['config']
    if cls == 'Sequential':
        recs = c['layers'] if isinstance(c, dict) else c          # Keras < 2.2.3 stored a bare list
        m = Sequential(name=c.get('name') if isinstance(c, dict) else None)
        for r in recs:
            if r['class_name'] == 'InputLayer':
                continue
            m.add(layer_from_config(r['class_name'], r['config'], reg))
        return m
    if cls != 'Model':
        raise ValueError('Unknown model class: ' + str(cls))
    layers, tensors = {}, {}
    pending = []
    for r in c['layers']:
        if r['class_name'] == 'InputLayer':
            shp = r['config']['batch_input_shape']
            t = Input(batch_shape=tuple(shp), name=r['name'])
            layers[r['name']] = t._node.layer
            tensors[(r['name'], 0)] = t
        else:
            layers[r['name']] = layer_from_config(r['class_name'], r['config'], reg)
            for k, node in enumerate(r['inbound_nodes']):
                pending.append((r['name'], k, node))
    # nodes may reference tensors created later in the file: resolve until nothing moves
    while pending:
        progressed = False
        for item in list(pending):
            name, k, node = item
            keys = [(i[0], i[1]) for i in node]
            if all(key in tensors for key in keys) and (k == 0 or (name, k - 1) in tensors):
                ins = [tensors[key] for key in keys]
                tensors[(name, k)] = layers[name](ins if len(ins) > 1 else ins[0])
                pending.remove(item)
                progressed = True
        if not progressed:
            raise ValueError('model_config graph cannot be resolved (cyclic or dangling inbound nodes)')
    ins = [tensors[(r[0], r[1])] for r in c['input_layers']]
    outs = [tensors[(r[0], r[1])] for r in c['output_layers']]
    return Model(inputs=ins[0] if len(ins) == 1 else ins, outputs=outs[0] if len(outs) == 1 else outs, name=c.get('name'))


# ---------------------------------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------------------------------

def _weight_names(layer):
    from .layers import ConvLSTM2D
    n = len(layer._weights)
    if n == 0:
        return []
    if isinstance(layer, ConvLSTM2D):
        names = ['kernel', 'recurrent_kernel', 'bias']
    else:
        names = ['kernel', 'bias']
    return ['%s/%s:0' % (layer.name, w) for w in names[:n]]


def _unique_layers(model):
    from .engine import InputLayer
    out = list(model.layers)
    if not any(isinstance(l, InputLayer) for l in out) and model.inputs:
        pass    # keras.Sequential.layers does not list the implicit InputLayer; neither does its weights group
    return out


def save_weights_to_group(model, g):
    layers = _unique_layers(model)
    g.attrs['layer_names'] = [l.name.encode('utf8') for l in layers]
    g.attrs['backend'] = BACKEND
    g.attrs['keras_version'] = KERAS_VERSION
    for l in layers:
        lg = g.create_group(l.name)
        names = _weight_names(l)
        lg.attrs['weight_names'] = [n.encode('utf8') for n in names] if names else np.zeros((0,), np.float64)
        for n, w in zip(names, l._weights):
            lg.create_dataset(n, np.asarray(w, np.float32))


def load_weights_from_group(model, g):
    names = [n.decode('utf8') if isinstance(n, bytes) else str(n) for n in np.atleast_1d(g.attrs['layer_names'])]
    stored = []
    for name in names:
        lg = g[name]
        wn = lg.attrs.get('weight_names')
        wn = [] if wn is None else [n.decode('utf8') if isinstance(n, bytes) else str(n) for n in np.atleast_1d(wn)
                                    if getattr(wn, 'dtype', None) is None or wn.dtype.kind in 'SOU']
        if wn:
            stored.append((name, [lg[n].read() for n in wn]))
    mine = [l for l in _unique_layers(model) if l._weights]
    if len(stored) != len(mine):
        raise ValueError('You are trying to load a weight file containing %d layers into a model with %d layers.' %
                         (len(stored), len(mine)))
    by_name = dict(stored)
    for k, l in enumerate(mine):
        ws = by_name[l.name] if l.name in by_name and len(by_name) == len(stored) else stored[k][1]
        l.set_weights(ws)


# ---------------------------------------------------------------------------------------------------------------------
# whole models
# ---------------------------------------------------------------------------------------------------------------------

def _loss_name(loss):
    if loss is None or isinstance(loss, str):
        return loss
    if isinstance(loss, (list, tuple)):
        return [_loss_name(v) for v in loss]
    if isinstance(loss, dict):
        return {k: _loss_name(v) for k, v in loss.items()}
    return getattr(loss, '__name__', loss.__class__.__name__)


def save_model(model, filepath, include_optimizer=True):
    w = hdf5.FileWriter()
    w.attrs['keras_version'] = KERAS_VERSION
    w.attrs['backend'] = BACKEND
    w.attrs['model_config'] = json.dumps(model_config(model)).encode('utf8')
    if include_optimizer and model.optimizer is not None:
        opt = model.optimizer
        ocfg = {k: _jsonable(v) for k, v in vars(opt).items() if isinstance(v, (int, float, bool, np.number))
                and k != 'iterations'}
        w.attrs['training_config'] = json.dumps({
            'optimizer_config': {'class_name': opt.__class__.__name__, 'config': ocfg},
            'loss': _loss_name(model.loss), 'metrics': _jsonable(model.metrics or []), 'sample_weight_mode': None,
            'loss_weights': _jsonable(model.loss_weights)}).encode('utf8')
    save_weights_to_group(model, w.create_group('model_weights'))
    if include_optimizer and model.loss is not None and _loss_name(model.loss) != model.loss:
        # Keras stores a custom loss by NAME and needs it back through custom_objects (DLWP/util.py:171-173); the picklable
        # loss objects of dlwp_b200.custom additionally travel with the file, in a group stock Keras ignores
        try:
            blob = pickle.dumps(model.loss, protocol=pickle.HIGHEST_PROTOCOL)
            w.create_dataset('dlwp_b200/loss_pickle', np.frombuffer(blob, np.uint8))
        except Exception:      # an unpicklable closure: by name only, like Keras
            pass
    w.save(filepath)


def _text(v):
    if isinstance(v, np.ndarray):
        v = v[()] if v.shape == () else v.ravel()[0]
    return v.decode('utf8') if isinstance(v, (bytes, np.bytes_)) else str(v)


def load_model(filepath, custom_objects=None, compile=True):
    from . import optimizers
    f = hdf5.File(filepath)
    if 'model_config' not in f.attrs:
        raise ValueError('No model found in config file.')
    model = model_from_config(json.loads(_text(f.attrs['model_config'])), custom_objects)
    load_weights_from_group(model, f['model_weights'])
    if compile and 'training_config' in f.attrs:
        tc = json.loads(_text(f.attrs['training_config']))
        oc = tc.get('optimizer_config', {})
        ocls = getattr(optimizers, oc.get('class_name', 'Adam'), None)
        if ocls is None:
            warnings.warn('optimizer %r is not implemented; the model is left uncompiled' % oc.get('class_name'))
            return model
        okw = {k: v for k, v in oc.get('config', {}).items()
               if k in inspect.signature(ocls.__init__).parameters}
        loss = tc.get('loss')
        objs = dict(custom_objects or {})
        if 'dlwp_b200' in f.keys() and 'loss_pickle' in f['dlwp_b200'].keys():
            stored = pickle.loads(f['dlwp_b200/loss_pickle'].read().tobytes())
            if _loss_name(stored) == loss:
                loss = stored

        def resolve(v):
            if isinstance(v, str) and v in objs:
                return objs[v]
            if isinstance(v, str) and v not in ('mse', 'mean_squared_error', 'mae', 'mean_absolute_error'):
                warnings.warn('loss %r is not in custom_objects; compiling with mean_squared_error' % v)
                return 'mse'
            return v
        if isinstance(loss, list):
            loss = [resolve(v) for v in loss]
        elif isinstance(loss, dict):
            loss = {k: resolve(v) for k, v in loss.items()}
        else:
            loss = resolve(loss)
        model.compile(optimizer=ocls(**okw), loss=loss, metrics=tc.get('metrics') or None,
                      loss_weights=tc.get('loss_weights'))
    return model
