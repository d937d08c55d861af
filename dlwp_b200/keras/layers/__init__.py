"""
keras.layers subset used by the DLWP convolutional example nets (examples/train.py:142-221,
examples/train_functional.py:154-285).  Each class only normalises its arguments, infers shapes and owns host copies of
its weights; the arithmetic is lowered to libdlwp_b200 by dlwp_b200.engine.
"""

import numpy as np

from ..engine import Input, InputLayer, KTensor, Layer, _glorot_uniform  # noqa: F401

_DEFAULT_DATA_FORMAT = 'channels_last'  # keras.json default (image_data_format)


def _norm_data_format(value):
    if value is None:
        return _DEFAULT_DATA_FORMAT
    v = str(value).lower()
    if v not in ('channels_first', 'channels_last'):
        raise ValueError('The `data_format` argument must be one of "channels_first", "channels_last". Received: ' +
                         str(value))
    return v


def _norm_tuple(value, n, name):
    if isinstance(value, (int, np.integer)):
        return (int(value),) * n
    try:
        t = tuple(int(v) for v in value)
    except (TypeError, ValueError):
        raise ValueError('The `%s` argument must be a tuple of %d integers. Received: %s' % (name, n, str(value)))
    if len(t) != n:
        raise ValueError('The `%s` argument must be a tuple of %d integers. Received: %s' % (name, n, str(value)))
    return t


def _norm_padding2d(padding):
    """keras ZeroPadding2D.__init__ argument normalisation (inherited by PeriodicPadding2D, DLWP/custom.py:183-189)."""
    if isinstance(padding, (int, np.integer)):
        return ((int(padding),) * 2,) * 2
    if hasattr(padding, '__len__'):
        if len(padding) != 2:
            raise ValueError('`padding` should have two elements. Found: ' + str(padding))
        return (_norm_tuple(padding[0], 2, '1st entry of padding'), _norm_tuple(padding[1], 2, '2nd entry of padding'))
    raise ValueError('`padding` should be either an int, a tuple of 2 ints (symmetric_height_pad, symmetric_width_pad),'
                     ' or a tuple of 2 tuples of 2 ints ((top_pad, bottom_pad), (left_pad, right_pad)). Found: ' +
                     str(padding))


class ZeroPadding2D(Layer):
    pad_mode = 'zero'

    def __init__(self, padding=(1, 1), data_format=None, **kwargs):
        super(ZeroPadding2D, self).__init__(**kwargs)
        self.data_format = _norm_data_format(data_format)
        self.padding = _norm_padding2d(padding)

    def compute_output_shape(self, s):
        (t, b), (l, r) = self.padding
        hi, wi = (2, 3) if self.data_format == 'channels_first' else (1, 2)
        out = list(s)
        out[hi] = None if s[hi] is None else s[hi] + t + b
        out[wi] = None if s[wi] is None else s[wi] + l + r
        return tuple(out)

    def get_config(self):
        c = super(ZeroPadding2D, self).get_config()
        c.update({'padding': self.padding, 'data_format': self.data_format})
        return c


class ZeroPadding3D(Layer):
    """keras ZeroPadding3D on (batch, depth, dim1, dim2, dim3) (channels_first) / (batch, dim1, dim2, dim3, depth).  The
    recurrent example nets feed it (batch, time, channels, lat, lon) tensors declared channels_first
    (examples/train.py:144-149), i.e. 'depth' is the time axis and dim1 the channel axis; the engine lowers the case
    that pads only the two trailing (lat, lon) axes -- a 2-D padding of every (time, channel) plane."""
    pad_mode = 'zero'

    def __init__(self, padding=(1, 1, 1), data_format=None, **kwargs):
        super(ZeroPadding3D, self).__init__(**kwargs)
        self.data_format = _norm_data_format(data_format)
        if isinstance(padding, (int, np.integer)):
            self.padding = ((int(padding),) * 2,) * 3
        else:
            if len(padding) != 3:
                raise ValueError('`padding` should have 3 elements. Found: ' + str(padding))
            self.padding = tuple(_norm_tuple(p, 2, 'entry of padding') for p in padding)

    def compute_output_shape(self, s):
        axes = (2, 3, 4) if self.data_format == 'channels_first' else (1, 2, 3)
        out = list(s)
        for a, (p0, p1) in zip(axes, self.padding):
            out[a] = None if s[a] is None else s[a] + p0 + p1
        return tuple(out)

    def get_config(self):
        c = super(ZeroPadding3D, self).get_config()
        c.update({'padding': self.padding, 'data_format': self.data_format})
        return c


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', data_format=None, dilation_rate=(1, 1),
                 activation=None, use_bias=True, kernel_initializer='glorot_uniform', bias_initializer='zeros',
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, **kwargs):
        super(Conv2D, self).__init__(**kwargs)
        self.filters = int(filters)
        self.kernel_size = _norm_tuple(kernel_size, 2, 'kernel_size')
        self.strides = _norm_tuple(strides, 2, 'strides')
        self.padding = str(padding).lower()
        if self.padding not in ('valid', 'same'):
            raise ValueError('The `padding` argument must be one of "valid", "same". Received: ' + str(padding))
        self.data_format = _norm_data_format(data_format)
        self.dilation_rate = _norm_tuple(dilation_rate, 2, 'dilation_rate')
        if activation is not None and not isinstance(activation, str):
            raise ValueError('dlwp_b200 Conv2D takes activation names (linear/tanh/relu), got %r' % (activation,))
        self.activation = activation
        self.use_bias = bool(use_bias)
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = bias_initializer
        self.kernel_regularizer = kernel_regularizer
        self.bias_regularizer = bias_regularizer

    def _channel_axis(self):
        return 1 if self.data_format == 'channels_first' else 3

    def build(self, s):
        cin = s[self._channel_axis()]
        if cin is None:
            raise ValueError('The channel dimension of the inputs should be defined. Found `None`.')
        kh, kw = self.kernel_size
        rng = np.random.RandomState(np.random.randint(0, 2 ** 31 - 1))
        kernel = _glorot_uniform(rng, (kh, kw, cin, self.filters), kh * kw * cin, kh * kw * self.filters)
        self._weights = [kernel] + ([np.zeros((self.filters,), np.float32)] if self.use_bias else [])
        self.built = True

    def _out_len(self, n, k, d, s):
        if n is None:
            return None
        if self.padding == 'same':
            return (n + s - 1) // s
        return (n - d * (k - 1) - 1) // s + 1

    def compute_output_shape(self, s):
        hi, wi = (2, 3) if self.data_format == 'channels_first' else (1, 2)
        out = list(s)
        out[hi] = self._out_len(s[hi], self.kernel_size[0], self.dilation_rate[0], self.strides[0])
        out[wi] = self._out_len(s[wi], self.kernel_size[1], self.dilation_rate[1], self.strides[1])
        out[self._channel_axis()] = self.filters
        for v in (out[hi], out[wi]):
            if v is not None and v <= 0:
                raise ValueError('Negative dimension size caused by the convolution of layer %s: input %s' %
                                 (self.name, s))
        return tuple(out)

    @property
    def kernel(self):
        return self._weights[0]

    @property
    def bias(self):
        return self._weights[1] if self.use_bias else None

    def get_config(self):
        c = super(Conv2D, self).get_config()
        c.update({'filters': self.filters, 'kernel_size': self.kernel_size, 'strides': self.strides,
                  'padding': self.padding, 'data_format': self.data_format, 'dilation_rate': self.dilation_rate,
                  'activation': self.activation, 'use_bias': self.use_bias})
        return c


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding='valid', data_format=None, **kwargs):
        super(MaxPooling2D, self).__init__(**kwargs)
        self.pool_size = _norm_tuple(pool_size, 2, 'pool_size')
        self.strides = self.pool_size if strides is None else _norm_tuple(strides, 2, 'strides')
        self.padding = str(padding).lower()
        self.data_format = _norm_data_format(data_format)

    def compute_output_shape(self, s):
        hi, wi = (2, 3) if self.data_format == 'channels_first' else (1, 2)
        out = list(s)
        for a, k, st in ((hi, self.pool_size[0], self.strides[0]), (wi, self.pool_size[1], self.strides[1])):
            if s[a] is not None:
                out[a] = (s[a] - k) // st + 1 if self.padding == 'valid' else (s[a] + st - 1) // st
        return tuple(out)

    def get_config(self):
        c = super(MaxPooling2D, self).get_config()
        c.update({'pool_size': self.pool_size, 'strides': self.strides, 'padding': self.padding,
                  'data_format': self.data_format})
        return c


class UpSampling2D(Layer):
    def __init__(self, size=(2, 2), data_format=None, interpolation='nearest', **kwargs):
        super(UpSampling2D, self).__init__(**kwargs)
        self.size = _norm_tuple(size, 2, 'size')
        self.data_format = _norm_data_format(data_format)
        if interpolation != 'nearest':
            raise ValueError('dlwp_b200 UpSampling2D implements interpolation="nearest" (the Keras default) only')
        self.interpolation = interpolation

    def compute_output_shape(self, s):
        hi, wi = (2, 3) if self.data_format == 'channels_first' else (1, 2)
        out = list(s)
        out[hi] = None if s[hi] is None else s[hi] * self.size[0]
        out[wi] = None if s[wi] is None else s[wi] * self.size[1]
        return tuple(out)

    def get_config(self):
        c = super(UpSampling2D, self).get_config()
        c.update({'size': self.size, 'data_format': self.data_format})
        return c


class Reshape(Layer):
    def __init__(self, target_shape, **kwargs):
        super(Reshape, self).__init__(**kwargs)
        self.target_shape = tuple(int(v) for v in target_shape)

    def compute_output_shape(self, s):
        known = int(np.prod([v for v in s[1:]]))
        tgt = list(self.target_shape)
        if tgt.count(-1) > 1:
            raise ValueError('Can only specify one unknown dimension.')
        if -1 in tgt:
            rest = int(np.prod([v for v in tgt if v != -1]))
            tgt[tgt.index(-1)] = known // max(rest, 1)
        if int(np.prod(tgt)) != known:
            raise ValueError('total size of new array must be unchanged')
        return (s[0],) + tuple(tgt)

    def get_config(self):
        c = super(Reshape, self).get_config()
        c.update({'target_shape': self.target_shape})
        return c


class Lambda(Layer):
    """Arbitrary host functions cannot run inside the GPU plan; DLWP.custom.slice_layer returns the `ChannelSlice`
    subclass below, which can (as a zero-copy channel window)."""

    def __init__(self, function, output_shape=None, mask=None, arguments=None, **kwargs):
        super(Lambda, self).__init__(**kwargs)
        self.function = function
        self.arguments = arguments or {}
        self._output_shape = output_shape

    def compute_output_shape(self, s):
        if self._output_shape is not None:
            shp = self._output_shape(s) if callable(self._output_shape) else (s[0],) + tuple(self._output_shape)
            return tuple(shp)
        probe = np.zeros([1 if v is None else v for v in s], np.float32)
        out = self.function(probe, **self.arguments)
        return (s[0],) + tuple(out.shape[1:])


class ChannelSlice(Lambda):
    """Lambda(x -> x[(slice(None),)*axis + (slice(start, end, step),)]) -- DLWP/custom.py:675-692."""

    def __init__(self, start, end, step=None, axis=1, **kwargs):
        self.start, self.end, self.step, self.axis = start, end, step, axis
        super(ChannelSlice, self).__init__(_SliceFunction(start, end, step, axis), **kwargs)

    def compute_output_shape(self, s):
        out = list(s)
        if s[self.axis] is not None:
            out[self.axis] = len(range(*slice(self.start, self.end, self.step).indices(s[self.axis])))
        return tuple(out)


class _SliceFunction(object):  # picklable stand-in for the closure in custom.py:685-690
    def __init__(self, start, end, step, axis):
        self.start, self.end, self.step, self.axis = start, end, step, axis

    def __call__(self, x):
        return x[tuple([slice(None)] * self.axis + [slice(self.start, self.end, self.step)])]


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super(Concatenate, self).__init__(**kwargs)
        self.axis = axis

    def compute_output_shape(self, shapes):
        if not isinstance(shapes, list) or len(shapes) < 2:
            raise ValueError('A `Concatenate` layer should be called on a list of at least 2 inputs')
        rank = len(shapes[0])
        ax = self.axis % rank
        out = list(shapes[0])
        for s in shapes[1:]:
            if len(s) != rank or any(a != b for k, (a, b) in enumerate(zip(s, shapes[0])) if k != ax):
                raise ValueError('A `Concatenate` layer requires inputs with matching shapes except for the concat '
                                 'axis. Got inputs shapes: %s' % shapes)
            out[ax] = None if (out[ax] is None or s[ax] is None) else out[ax] + s[ax]
        return tuple(out)

    def get_config(self):
        c = super(Concatenate, self).get_config()
        c.update({'axis': self.axis})
        return c


def concatenate(inputs, axis=-1, **kwargs):
    return Concatenate(axis=axis, **kwargs)(inputs)


class _NotOnHotPath(Layer):
    """Importable so that `from keras.layers import ...` lines of the example scripts work
    (examples/train_functional.py:21-22); building one raises -- dense / locally connected nets are not on the hot path."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError('%s is not part of the convolutional rollout hot path implemented by dlwp_b200 '
                                  '(SURVEY.md section 8f)' % self.__class__.__name__)


class _LstmPart(object):
    """One of ConvLSTM2D's two convolutions as the engine's weight table sees it (a Conv2D-like weight owner)."""

    def __init__(self, layer, recurrent):
        self.layer, self.recurrent = layer, recurrent
        self.name = layer.name + ('/recurrent_kernel' if recurrent else '/kernel')

    @property
    def _weights(self):
        w = self.layer._weights
        return [w[1]] if self.recurrent else [w[0]] + ([w[2]] if self.layer.use_bias else [])

    @property
    def use_bias(self):
        return self.layer.use_bias and not self.recurrent

    @property
    def _weights_version(self):
        return self.layer._weights_version

    @property
    def kernel_regularizer(self):
        return self.layer.recurrent_regularizer if self.recurrent else self.layer.kernel_regularizer

    @property
    def bias_regularizer(self):
        return None if self.recurrent else self.layer.bias_regularizer


class ConvLSTM2D(Layer):
    """keras.layers.ConvLSTM2D (Keras 2.2 signature) as the recurrent front block of the example nets uses it
    (examples/train.py:144-157, examples/train_functional.py:155-166): input (batch, time, channels, rows, cols) for
    channels_first.  Weights in Keras order and layout: kernel (kh, kw, Cin, 4F), recurrent_kernel (kh, kw, F, 4F), bias
    (4F,), gate blocks i, f, c, o; `unit_forget_bias` initialises the f block of the bias with ones.  The input convolution
    uses the layer's padding / dilation, the recurrent one is always 'same', undilated (keras ConvLSTM2DCell)."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', data_format=None, dilation_rate=(1, 1),
                 activation='tanh', recurrent_activation='hard_sigmoid', use_bias=True,
                 kernel_initializer='glorot_uniform', recurrent_initializer='orthogonal', bias_initializer='zeros',
                 unit_forget_bias=True, kernel_regularizer=None, recurrent_regularizer=None, bias_regularizer=None,
                 activity_regularizer=None, kernel_constraint=None, recurrent_constraint=None, bias_constraint=None,
                 return_sequences=False, go_backwards=False, stateful=False, dropout=0., recurrent_dropout=0., **kwargs):
        super(ConvLSTM2D, self).__init__(**kwargs)
        self.filters = int(filters)
        self.kernel_size = _norm_tuple(kernel_size, 2, 'kernel_size')
        self.strides = _norm_tuple(strides, 2, 'strides')
        self.padding = str(padding).lower()
        if self.padding not in ('valid', 'same'):
            raise ValueError('The `padding` argument must be one of "valid", "same". Received: ' + str(padding))
        self.data_format = _norm_data_format(data_format)
        self.dilation_rate = _norm_tuple(dilation_rate, 2, 'dilation_rate')
        for a in (activation, recurrent_activation):
            if a is not None and not isinstance(a, str):
                raise ValueError('dlwp_b200 ConvLSTM2D takes activation names, got %r' % (a,))
        self.activation = activation
        self.recurrent_activation = recurrent_activation
        self.use_bias = bool(use_bias)
        self.unit_forget_bias = bool(unit_forget_bias)
        self.kernel_regularizer = kernel_regularizer
        self.recurrent_regularizer = recurrent_regularizer
        self.bias_regularizer = bias_regularizer
        self.return_sequences = bool(return_sequences)
        self.go_backwards = bool(go_backwards)
        self.stateful = bool(stateful)
        self.dropout, self.recurrent_dropout = float(dropout), float(recurrent_dropout)
        self._parts = (_LstmPart(self, False), _LstmPart(self, True))

    def _axes(self):   # (channel, rows, cols) of the 5-D input
        return (2, 3, 4) if self.data_format == 'channels_first' else (4, 2, 3)

    def build(self, s):
        if len(s) != 5:
            raise ValueError('Input 0 is incompatible with layer %s: expected ndim=5, found ndim=%d' % (self.name, len(s)))
        cin = s[self._axes()[0]]
        if cin is None:
            raise ValueError('The channel dimension of the inputs should be defined. Found `None`.')
        kh, kw = self.kernel_size
        F = self.filters
        rng = np.random.RandomState(np.random.randint(0, 2 ** 31 - 1))
        kernel = _glorot_uniform(rng, (kh, kw, cin, 4 * F), kh * kw * cin, kh * kw * 4 * F)
        # keras' default recurrent initializer is 'orthogonal': Q of a QR of a normal (kh*kw*F, 4F) matrix
        a = rng.standard_normal((kh * kw * F, 4 * F))
        q, r = np.linalg.qr(a if a.shape[0] >= a.shape[1] else a.T)
        q = q * np.sign(np.diag(r))
        rec = (q if a.shape[0] >= a.shape[1] else q.T).reshape(kh, kw, F, 4 * F).astype(np.float32)
        self._weights = [kernel, rec]
        if self.use_bias:
            b = np.zeros((4 * F,), np.float32)
            if self.unit_forget_bias:
                b[F:2 * F] = 1.0
            self._weights.append(b)
        self.built = True

    def compute_output_shape(self, s):
        ca, ha, wa = self._axes()

        def out_len(n, k, d, st):
            if n is None:
                return None
            return (n + st - 1) // st if self.padding == 'same' else (n - d * (k - 1) - 1) // st + 1
        rows = out_len(s[ha], self.kernel_size[0], self.dilation_rate[0], self.strides[0])
        cols = out_len(s[wa], self.kernel_size[1], self.dilation_rate[1], self.strides[1])
        for v in (rows, cols):
            if v is not None and v <= 0:
                raise ValueError('Negative dimension size caused by the convolution of layer %s: input %s' %
                                 (self.name, s))
        if self.data_format == 'channels_first':
            tail = (self.filters, rows, cols)
        else:
            tail = (rows, cols, self.filters)
        return (s[0], s[1]) + tail if self.return_sequences else (s[0],) + tail

    @property
    def kernel(self):
        return self._weights[0]

    @property
    def recurrent_kernel(self):
        return self._weights[1]

    @property
    def bias(self):
        return self._weights[2] if self.use_bias else None

    def __getstate__(self):
        return self.__dict__.copy()

    def get_config(self):
        c = super(ConvLSTM2D, self).get_config()
        c.update({'filters': self.filters, 'kernel_size': self.kernel_size, 'strides': self.strides,
                  'padding': self.padding, 'data_format': self.data_format, 'dilation_rate': self.dilation_rate,
                  'activation': self.activation, 'recurrent_activation': self.recurrent_activation,
                  'use_bias': self.use_bias, 'unit_forget_bias': self.unit_forget_bias,
                  'return_sequences': self.return_sequences, 'go_backwards': self.go_backwards,
                  'stateful': self.stateful, 'dropout': self.dropout, 'recurrent_dropout': self.recurrent_dropout})
        return c


class LocallyConnected2D(_NotOnHotPath):
    pass


class Dense(_NotOnHotPath):
    pass


class Flatten(_NotOnHotPath):
    pass


class BatchNormalization(_NotOnHotPath):
    pass
