"""
Core of the minimal Keras-2.2-compatible front end: symbolic tensors, Layer, Model, Sequential.

Only graph *construction* and host-side bookkeeping live here (what the reference delegates to `keras.layers` /
`keras.models`, DLWP/model/models.py:96-112); every FLOP of `Model.predict` runs in libdlwp_b200.so through
dlwp_b200.engine.CompiledNet.  Layer names follow Keras' auto-naming (`conv2d_1`, `periodic_padding2d_1`, ...), shapes
carry the leading `None` batch axis, weights use Keras layouts -- so the example scripts' introspection
(`model.layers[i].name / .output_shape`, examples/train.py:242-243) reads the same.
"""

import pickle
import re

import sys

import numpy as np

_NAME_COUNTS = {}


def _snake(name):
    s = re.sub('(.)([A-Z][a-z0-9]+)', r'\1_\2', name)
    return re.sub('([a-z])([A-Z])', r'\1_\2', s).lower()


def _auto_name(cls_name):
    base = _snake(cls_name)
    _NAME_COUNTS[base] = _NAME_COUNTS.get(base, 0) + 1
    return '%s_%d' % (base, _NAME_COUNTS[base])


def clear_session():
    """keras.backend.clear_session(): resets the auto-naming counters (examples/validate.py:245)."""
    _NAME_COUNTS.clear()


class KTensor(object):
    """Symbolic tensor: a shape (with leading None) and the node that produces it."""

    def __init__(self, shape, node=None, index=0):
        self.shape = tuple(shape)
        self._node = node
        self._index = index
        self._keras_shape = self.shape

    def __repr__(self):
        return '<KTensor shape=%s from %s>' % (self.shape, self._node.layer.name if self._node else None)


class Node(object):
    def __init__(self, layer, inputs):
        self.layer = layer
        self.inputs = list(inputs)
        self.output = None


class Layer(object):
    def __init__(self, name=None, input_shape=None, batch_input_shape=None, trainable=True, dtype=None, **kwargs):
        if kwargs:
            raise TypeError('Keyword argument not understood: %s' % sorted(kwargs)[0])
        self.name = name or _auto_name(self.__class__.__name__)
        if batch_input_shape is None and input_shape is not None:
            batch_input_shape = (None,) + tuple(input_shape)
        self._batch_input_shape = batch_input_shape
        self.trainable = trainable
        self.built = False
        self._weights = []
        self._weights_version = 0
        self._inbound_nodes = []

    # -- to be specialised ---------------------------------------------------------------------------------------
    def build(self, input_shape):
        self.built = True

    def compute_output_shape(self, input_shape):
        return input_shape

    def get_config(self):
        return {'name': self.name, 'trainable': self.trainable}

    # -- graph construction --------------------------------------------------------------------------------------
    def __call__(self, inputs):
        many = isinstance(inputs, (list, tuple))
        ins = list(inputs) if many else [inputs]
        for t in ins:
            if not isinstance(t, KTensor):
                raise ValueError('Layer %s was called with an input that isn\'t a symbolic tensor: %r' %
                                 (self.name, type(t)))
        in_shape = [t.shape for t in ins] if many else ins[0].shape
        if not self.built:
            self.build(in_shape)
            self.built = True
        node = Node(self, ins)
        node.output = KTensor(self.compute_output_shape(in_shape), node)
        self._inbound_nodes.append(node)
        return node.output

    @property
    def input_shape(self):
        if not self._inbound_nodes:
            raise AttributeError('The layer has never been called and thus has no defined input shape.')
        ins = self._inbound_nodes[0].inputs
        return ins[0].shape if len(ins) == 1 else [t.shape for t in ins]

    @property
    def output_shape(self):
        if not self._inbound_nodes:
            raise AttributeError('The layer has never been called and thus has no defined output shape.')
        return self._inbound_nodes[0].output.shape

    @property
    def output(self):
        return self._inbound_nodes[0].output

    # -- weights -------------------------------------------------------------------------------------------------
    @property
    def weights(self):
        return list(self._weights)

    def get_weights(self):
        return [w.copy() for w in self._weights]

    def set_weights(self, weights):
        weights = list(weights)
        if len(weights) != len(self._weights):
            raise ValueError('You called `set_weights(weights)` on layer "%s" with a weight list of length %d, but '
                             'the layer was expecting %d weights.' % (self.name, len(weights), len(self._weights)))
        new = []
        for w, old in zip(weights, self._weights):
            w = np.asarray(w, dtype=np.float32)
            if w.shape != old.shape:
                raise ValueError('Layer weight shape %s not compatible with provided weight shape %s' %
                                 (old.shape, w.shape))
            new.append(np.ascontiguousarray(w))
        self._weights = new
        self._weights_version += 1

    def count_params(self):
        return int(sum(w.size for w in self._weights))

    def __getstate__(self):
        return self.__dict__.copy()


class InputLayer(Layer):
    def __init__(self, input_shape=None, batch_input_shape=None, name=None, **kwargs):
        super(InputLayer, self).__init__(name=name or _auto_name('input'), input_shape=input_shape,
                                         batch_input_shape=batch_input_shape, **kwargs)
        self.built = True
        node = Node(self, [])
        node.output = KTensor(self._batch_input_shape, node)
        self._inbound_nodes.append(node)

    @property
    def input_shape(self):
        return self._batch_input_shape


def Input(shape=None, batch_shape=None, name=None, dtype=None, **kwargs):
    if shape is None and batch_shape is None:
        raise ValueError('Please provide to Input either a `shape` or a `batch_shape` argument.')
    return InputLayer(input_shape=shape, batch_input_shape=batch_shape, name=name).output


# ---------------------------------------------------------------------------------------------------------------- #
# Models
# ---------------------------------------------------------------------------------------------------------------- #

def _glorot_uniform(rng, shape, fan_in, fan_out):
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


class History(object):
    """keras.callbacks.History."""

    def __init__(self):
        self.epoch = []
        self.history = {}
        self.model = None
        self.params = {}

    def set_model(self, model):
        self.model = model

    def set_params(self, params):
        self.params = params

    def on_train_begin(self, logs=None):
        self.epoch = []
        self.history = {}

    def on_train_end(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        self.epoch.append(epoch)
        for k, v in (logs or {}).items():
            self.history.setdefault(k, []).append(v)

    def on_batch_begin(self, batch, logs=None):
        pass

    def on_batch_end(self, batch, logs=None):
        pass


class Model(object):
    """keras.models.Model over the symbolic graph; `predict` executes on the GPU through libdlwp_b200."""

    def __init__(self, inputs=None, outputs=None, name=None):
        self.name = name or _auto_name('model')
        self.stop_training = False
        self.optimizer = None
        self.loss = None
        self.loss_weights = None
        self.metrics = None
        self._compile_kwargs = None
        self._engine = None
        self._engine_key = None
        self.history = None
        if inputs is not None:
            self._init_graph(inputs, outputs)

    def _init_graph(self, inputs, outputs):
        self._single_output = not isinstance(outputs, (list, tuple))
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = [outputs] if self._single_output else list(outputs)
        if len(self.inputs) != 1:
            raise NotImplementedError('dlwp_b200 supports single-input models (all DLWP example nets are)')
        order, seen = [], set()

        def visit(t):
            node = t._node
            if node is None:
                raise ValueError('Graph disconnected: tensor %r has no producer' % (t,))
            if id(node) in seen:
                return
            seen.add(id(node))
            for s in node.inputs:
                visit(s)
            order.append(node)
        import sys
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
        for o in self.outputs:
            visit(o)
        if id(self.inputs[0]._node) not in seen:
            raise ValueError('Graph disconnected: the model input does not reach the outputs')
        self._nodes = order
        self.layers = []
        for node in order:
            if node.layer not in self.layers:
                self.layers.append(node.layer)
        self._engine = None

    # -- introspection -------------------------------------------------------------------------------------------
    @property
    def input_shape(self):
        return self.inputs[0].shape

    @property
    def output_shape(self):
        return self.outputs[0].shape if len(self.outputs) == 1 else [o.shape for o in self.outputs]

    def get_layer(self, name=None, index=None):
        if index is not None:
            return self.layers[index]
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError('No such layer: ' + str(name))

    def count_params(self):
        return int(sum(l.count_params() for l in self.layers))

    def summary(self, line_length=None, positions=None, print_fn=None):
        print_fn = print_fn or print
        print_fn('_' * 65)
        print_fn('%-29s%-26s%-10s' % ('Layer (type)', 'Output Shape', 'Param #'))
        print_fn('=' * 65)
        for l in self.layers:
            print_fn('%-29s%-26s%-10d' % ('%s (%s)' % (l.name, l.__class__.__name__), str(l.output_shape),
                                          l.count_params()))
        print_fn('=' * 65)
        print_fn('Total params: {:,}'.format(self.count_params()))
        print_fn('_' * 65)

    # -- weights -------------------------------------------------------------------------------------------------
    def get_weights(self):
        return [w for l in self.layers for w in l.get_weights()]

    def set_weights(self, weights):
        weights = list(weights)
        n = sum(len(l._weights) for l in self.layers)
        if len(weights) != n:
            raise ValueError('You called `set_weights(weights)` on model "%s" with a weight list of length %d, but '
                             'the model was expecting %d weights.' % (self.name, len(weights), n))
        for l in self.layers:
            k = len(l._weights)
            if k:
                l.set_weights(weights[:k])
                weights = weights[k:]

    def reset_states(self):
        pass  # no stateful layers on the convolutional hot path (custom.RNNResetStates calls this every epoch)

    # -- compile / predict ---------------------------------------------------------------------------------------
    def compile(self, optimizer=None, loss=None, metrics=None, loss_weights=None, **kwargs):
        from . import optimizers
        self.optimizer = optimizers.get(optimizer) if optimizer is not None else None
        self.loss = loss
        self.metrics = metrics or []
        self.loss_weights = loss_weights
        self._compile_kwargs = dict(optimizer=optimizer, loss=loss, metrics=metrics, loss_weights=loss_weights,
                                    **kwargs)
        stale = self.__dict__.pop('_train_engine', None)   # a new loss / optimizer: the training plan (its loss kind, Adam
        if stale is not None:                              # moments and step count) starts afresh, as in Keras
            stale.close()

    def engine(self, batch=None, device=None):
        """The compiled GPU plan (built lazily; rebuilt when a larger batch chunk is needed).  `device`: CUDA device index
        of a replica of a multi_gpu_model (None: the current device)."""
        from ..engine import CompiledNet
        if device is not None:
            import torch
            reps = self.__dict__.setdefault('_replicas', {})
            eng = reps.get(device)
            if eng is None or (batch is not None and not eng.fits(batch)):
                with torch.cuda.device(device):
                    if eng is not None:
                        eng.close()
                    eng = reps[device] = CompiledNet(self, batch or 1)
            return eng
        if self._engine is None or (batch is not None and not self._engine.fits(batch)):
            if self._engine is not None:
                self._engine.close()
            self._engine = CompiledNet(self, batch or 1)
        return self._engine

    # -- keras.utils.multi_gpu_model (DLWP/model/models.py:104-109): single-process batch split over `gpus` devices ------
    def _split(self, x, call, axis):
        """Run call(engine, chunk) for the slices of x dealt to the replicas (one host thread per device: the library calls
        and the copies release the GIL) and concatenate the results along `axis` of every returned array."""
        import torch
        from concurrent.futures import ThreadPoolExecutor
        gpus = int(getattr(self, '_gpus', 1) or 1)
        if gpus > torch.cuda.device_count():
            raise ValueError('To call `multi_gpu_model` with `gpus=%d`, we expect the following devices to be available: '
                             '%s. However this machine only has: %d CUDA device(s).' %
                             (gpus, ['/gpu:%d' % i for i in range(gpus)], torch.cuda.device_count()))
        n = x.shape[0]
        bounds = [(n * d) // gpus for d in range(gpus + 1)]
        parts = [(d, x[bounds[d]:bounds[d + 1]]) for d in range(gpus) if bounds[d + 1] > bounds[d]]

        def work(item):
            d, chunk = item
            with torch.cuda.device(d):
                return call(self.engine(chunk.shape[0], device=d), chunk)
        with ThreadPoolExecutor(max_workers=len(parts)) as pool:
            res = list(pool.map(work, parts))
        if isinstance(res[0], (list, tuple)):
            return [np.concatenate([r[k] for r in res], axis=axis) for k in range(len(res[0]))]
        return np.concatenate(res, axis=axis)

    def _multi(self):
        return int(getattr(self, '_gpus', 1) or 1) > 1

    def rollout_host(self, x, steps):
        """Device-resident rollout, numpy in / numpy (slots, N, ...) out; a multi_gpu_model splits the batch over devices."""
        if self._multi():
            return self._split(np.asarray(x), lambda eng, c: eng.rollout_host(c, steps), axis=1)
        return self.engine(len(x)).rollout_host(x, steps)

    def rollout_step_sequence_host(self, x, steps, time_dim):
        if self._multi():
            return self._split(np.asarray(x), lambda eng, c: eng.rollout_step_sequence_host(c, steps, time_dim), axis=1)
        return self.engine(len(x)).rollout_step_sequence_host(x, steps, time_dim)

    def predict(self, x, batch_size=None, verbose=0, steps=None, **kwargs):
        """keras.Model.predict: numpy in, numpy (or list of numpy) out.  `batch_size` only affects Keras' host-side
        chunking, which has no numerical effect; the GPU plan picks its own chunk size."""
        x = np.asarray(x)
        exp = self.inputs[0].shape[1:]
        if x.shape[1:] != tuple(exp):
            raise ValueError('Error when checking input: expected %s to have shape %s but got array with shape %s' %
                             (self.inputs[0]._node.layer.name, exp, x.shape[1:]))
        if self._multi():
            outs = self._split(x, lambda eng, c: eng.predict(c), axis=0)
        else:
            outs = self.engine(x.shape[0]).predict(x)
        return outs[0] if self._single_output_list() else outs

    def _single_output_list(self):
        return len(self.outputs) == 1

    def predict_on_batch(self, x):
        return self.predict(x)

    def fit(self, *args, **kwargs):
        from ..training import fit
        return fit(self, *args, **kwargs)

    def fit_generator(self, *args, **kwargs):
        from ..training import fit_generator
        return fit_generator(self, *args, **kwargs)

    def evaluate(self, *args, **kwargs):
        from ..training import evaluate
        return evaluate(self, *args, **kwargs)

    def train_on_batch(self, *args, **kwargs):
        from ..training import train_on_batch
        return train_on_batch(self, *args, **kwargs)

    # -- persistence ---------------------------------------------------------------------------------------------
    def __getstate__(self):
        # runtime handles (ctypes plan pointers, torch modules, the History callback's model back-reference) never travel
        d = self.__dict__.copy()
        for k in ('_engine', '_engine_key', '_train_engine', '_train_key', 'history'):
            if k in d:
                d[k] = None
        d.pop('_replicas', None)
        return d

    def save(self, filepath, overwrite=True, include_optimizer=True):
        """keras.Model.save: the Keras 2.2 HDF5 model file (model_config + training_config + model_weights group), written
        by dlwp_b200.hdf5 -- the format DLWP/util.py:139-141 produces and :173 reads (keras/saving.py has the layout)."""
        from .saving import save_model
        save_model(self, filepath, include_optimizer=include_optimizer)

    def save_weights(self, filepath):
        """keras.Model.save_weights: an HDF5 file whose root is the `model_weights` group of a full model file."""
        from .. import hdf5
        from .saving import save_weights_to_group
        w = hdf5.FileWriter()
        save_weights_to_group(self, w)
        w.save(filepath)

    def load_weights(self, filepath):
        from .. import hdf5
        from .saving import load_weights_from_group
        f = hdf5.File(filepath)
        load_weights_from_group(self, f['model_weights'] if 'model_weights' in f.keys() else f)

    def get_config(self):
        from .saving import model_config
        return model_config(self)['config']

    def to_json(self):
        import json
        from .saving import model_config
        return json.dumps(model_config(self))

    def __del__(self):
        if sys is None or sys.is_finalizing():
            return
        try:
            if self._engine is not None:
                self._engine.close()
            for eng in self.__dict__.get('_replicas', {}).values():
                eng.close()
        except Exception:
            pass


class Sequential(Model):
    def __init__(self, layers=None, name=None):
        super(Sequential, self).__init__(name=name or _auto_name('sequential'))
        self.layers = []
        self.inputs = []
        self.outputs = []
        self._single_output = True
        for l in layers or []:
            self.add(l)

    def add(self, layer):
        if not isinstance(layer, Layer):
            raise TypeError('The added layer must be an instance of class Layer. Found: ' + str(layer))
        if not self.outputs:
            if isinstance(layer, InputLayer):
                x = layer.output
                self.inputs = [x]
                self.outputs = [x]
                self._user_layers = []
                self._rebuild()
                return
            if layer._batch_input_shape is None:
                raise ValueError('The first layer in a Sequential model must get an `input_shape` or '
                                 '`batch_input_shape` argument.')
            x = InputLayer(batch_input_shape=layer._batch_input_shape).output
            self.inputs = [x]
            self.outputs = [layer(x)]
            self._user_layers = [layer]
        else:
            self.outputs = [layer(self.outputs[0])]
            self._user_layers.append(layer)
        self._rebuild()

    def _rebuild(self):
        self._init_graph(self.inputs[0], self.outputs[0])
        # keras.Sequential.layers does not list the implicit InputLayer
        self.layers = [l for l in self.layers if not isinstance(l, InputLayer)]


def load_model(filepath, custom_objects=None, compile=True):
    """keras.models.load_model for Keras 2.2 HDF5 model files (written by the reference or by Model.save above); the
    pickle container of this package's first release is still read."""
    with open(filepath, 'rb') as f:
        head = f.read(8)
    if head[:2] == b'\x80\x04' or head[:2] == b'\x80\x05' or head[:1] == b'\x80':
        with open(filepath, 'rb') as f:
            blob = pickle.load(f)
        if isinstance(blob, dict) and blob.get('format') == 'dlwp_b200.keras.v1':
            return blob['model']
        raise ValueError('%s is neither an HDF5 model file nor a dlwp_b200 pickle container' % filepath)
    from .saving import load_model as _load
    return _load(filepath, custom_objects=custom_objects, compile=compile)
