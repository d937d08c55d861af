"""
Layer objects, the (name, args, kwargs) layer-tuple interpreter and the functional example graphs.  TEST INFRASTRUCTURE.

* ``OSequential`` follows DLWPNeuralNet.build_model (DLWP/model/models.py:63-112): the same tuple validation and the
  same "look the name up, instantiate with *args/**kwargs, append" construction, over the layer names that the
  convolutional examples use (examples/train.py:159-221).
* ``OFunctionalNet`` follows examples/train_functional.py:154-285 (``basic_model`` 222-245, ``skip_model`` 248-275,
  the shared-weight unroll 278-281) for the non-recurrent case.

Every layer is callable on a numpy array (computes in the array's dtype: float64 = tier-0 oracle) or on a torch CPU
tensor (fp32 oneDNN = tier-1 "reference precision" and the CPU timing stand-in).
"""

import numpy as np

from . import ops

try:  # torch is only needed for the tier-1 / timing path
    import torch
    import torch.nn.functional as F
except ImportError:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


class OLayer(object):
    weights = ()

    def output_shape(self, s):
        return s

    def build(self, s):
        pass


class OPeriodicPadding2D(OLayer):
    """DLWP/custom.py:139-214."""

    def __init__(self, padding=(1, 1), data_format=None, **kwargs):
        self.padding = ops.normalize_padding(padding)
        self.data_format = data_format or 'channels_last'  # Keras default image_data_format

    def output_shape(self, s):
        (t, b), (l, r) = self.padding
        if self.data_format == 'channels_first':
            return (s[0], s[1] + t + b, s[2] + l + r)
        return (s[0] + t + b, s[1] + l + r, s[2])

    def __call__(self, x):
        if _is_torch(x):
            assert self.data_format == 'channels_first'
            (t, b), (l, r) = self.padding
            if l or r:
                x = torch.cat([x[..., x.shape[-1] - l:], x, x[..., :r]], dim=-1)
            if t or b:
                x = torch.cat([x[..., x.shape[-2] - t:, :], x, x[..., :b, :]], dim=-2)
            return x
        return ops.periodic_pad2d(x, self.padding, self.data_format)


class OZeroPadding2D(OPeriodicPadding2D):
    def __call__(self, x):
        if _is_torch(x):
            assert self.data_format == 'channels_first'
            (t, b), (l, r) = self.padding
            return F.pad(x, (l, r, t, b))
        return ops.zero_pad2d(x, self.padding, self.data_format)


class OFillPadding2D(OPeriodicPadding2D):
    """DLWP/custom.py:309-402 (numpy only)."""

    def __call__(self, x):
        return ops.fill_pad2d(np.asarray(x), self.padding, self.data_format)


class OTFPadding2D(OPeriodicPadding2D):
    """DLWP/custom.py:527-599 (numpy only)."""

    def __init__(self, padding=(1, 1), data_format=None, mode='CONSTANT', constant_values=0, **kwargs):
        super(OTFPadding2D, self).__init__(padding, data_format)
        if mode.upper() == 'CONSTANT' and constant_values != 0:
            raise NotImplementedError('constant_values != 0')
        self.mode = mode

    def __call__(self, x):
        return ops.tf_pad2d(np.asarray(x), self.padding, self.mode, self.data_format)


class OConv2D(OLayer):
    """Keras Conv2D, 'valid', cross-correlation, kernel (kh,kw,Cin,Cout) -- SURVEY.md Appendix A.2."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', data_format=None, dilation_rate=(1, 1),
                 activation=None, use_bias=True, **kwargs):
        if padding != 'valid':
            raise ValueError("oracle Conv2D only restates padding='valid' (the only mode on the hot path)")
        self.filters = int(filters)
        self.kernel_size = ops.normalize_pair(kernel_size, 'kernel_size')
        self.strides = ops.normalize_pair(strides, 'strides')
        self.dilation_rate = ops.normalize_pair(dilation_rate, 'dilation_rate')
        self.data_format = data_format or 'channels_last'
        self.activation = activation
        self.use_bias = use_bias
        self.kernel = None
        self.bias = None
        self._tw = None

    def build(self, s):
        cin = s[0] if self.data_format == 'channels_first' else s[2]
        self.kernel = np.zeros(self.kernel_size + (cin, self.filters), np.float32)
        self.bias = np.zeros((self.filters,), np.float32) if self.use_bias else None

    def output_shape(self, s):
        kh, kw = self.kernel_size
        dh, dw = self.dilation_rate
        sh, sw = self.strides
        if self.data_format == 'channels_first':
            return (self.filters, (s[1] - dh * (kh - 1) - 1) // sh + 1, (s[2] - dw * (kw - 1) - 1) // sw + 1)
        return ((s[0] - dh * (kh - 1) - 1) // sh + 1, (s[1] - dw * (kw - 1) - 1) // sw + 1, self.filters)

    @property
    def weights(self):
        return [self.kernel] + ([self.bias] if self.use_bias else [])

    def set_weights(self, ws):
        self.kernel = np.asarray(ws[0])
        if self.use_bias:
            self.bias = np.asarray(ws[1])
        self._tw = None

    def __call__(self, x):
        if _is_torch(x):
            assert self.data_format == 'channels_first'
            if getattr(self, '_tparam', None) is not None:   # autograd leaves supplied by a gradient test
                self._tw = self._tparam
            if self._tw is None:
                w = torch.from_numpy(np.ascontiguousarray(np.transpose(self.kernel, (3, 2, 0, 1)))).to(x.dtype)
                b = torch.from_numpy(np.ascontiguousarray(self.bias)).to(x.dtype) if self.use_bias else None
                self._tw = (w, b)
            y = F.conv2d(x, self._tw[0], self._tw[1], stride=self.strides, dilation=self.dilation_rate)
            if self.activation == 'tanh':
                y = torch.tanh(y)
            elif self.activation == 'relu':
                y = torch.relu(y)
            elif self.activation not in (None, 'linear'):
                raise ValueError(self.activation)
            return y
        k = self.kernel.astype(x.dtype)
        b = self.bias.astype(x.dtype) if self.use_bias else None
        y = ops.conv2d_valid(x, k, b, self.dilation_rate, self.strides, self.data_format)
        return ops.activation(self.activation)(y)


class ORowConnected2D(OConv2D):
    """DLWP/custom.py:695-837: kernel (H_out,kh,kw,Cin,Cout), bias (H_out,1,Cout)."""

    def build(self, s):
        cin = s[0] if self.data_format == 'channels_first' else s[2]
        ho = self.output_shape(s)[1 if self.data_format == 'channels_first' else 0]
        self.kernel = np.zeros((ho,) + self.kernel_size + (cin, self.filters), np.float32)
        self.bias = np.zeros((ho, 1, self.filters), np.float32) if self.use_bias else None

    def __call__(self, x):
        if _is_torch(x):
            y = torch.from_numpy(self(x.numpy()))
            return y
        k = self.kernel.astype(x.dtype)
        b = self.bias.astype(x.dtype) if self.use_bias else None
        y = ops.row_conv2d(x, k, b, self.strides, self.data_format)
        return ops.activation(self.activation)(y)


class OMaxPooling2D(OLayer):
    def __init__(self, pool_size=(2, 2), strides=None, padding='valid', data_format=None, **kwargs):
        self.pool_size = ops.normalize_pair(pool_size, 'pool_size')
        if strides is not None and ops.normalize_pair(strides) != self.pool_size:
            raise ValueError('oracle MaxPooling2D only restates strides == pool_size')
        self.data_format = data_format or 'channels_last'

    def output_shape(self, s):
        ph, pw = self.pool_size
        if self.data_format == 'channels_first':
            return (s[0], s[1] // ph, s[2] // pw)
        return (s[0] // ph, s[1] // pw, s[2])

    def __call__(self, x):
        if _is_torch(x):
            return F.max_pool2d(x, self.pool_size)
        return ops.max_pool2d(x, self.pool_size, self.data_format)


class OUpSampling2D(OLayer):
    def __init__(self, size=(2, 2), data_format=None, **kwargs):
        self.size = ops.normalize_pair(size, 'size')
        self.data_format = data_format or 'channels_last'

    def output_shape(self, s):
        sh, sw = self.size
        if self.data_format == 'channels_first':
            return (s[0], s[1] * sh, s[2] * sw)
        return (s[0] * sh, s[1] * sw, s[2])

    def __call__(self, x):
        if _is_torch(x):
            return x.repeat_interleave(self.size[0], dim=-2).repeat_interleave(self.size[1], dim=-1)
        return ops.upsample2d(x, self.size, self.data_format)


class OPeriodicPadding3D(OLayer):
    """DLWP/custom.py:217-306 on per-sample shapes (d, a1, a2, a3) (channels_first: the three trailing axes are padded)."""
    _pad = staticmethod(ops.periodic_pad3d)

    def __init__(self, padding=(1, 1, 1), data_format=None, **kwargs):
        self.padding = ops.normalize_padding3d(padding)
        self.data_format = data_format or 'channels_last'

    def output_shape(self, s):
        axes = (1, 2, 3) if self.data_format == 'channels_first' else (0, 1, 2)
        s = list(s)
        for a, (p0, p1) in zip(axes, self.padding):
            s[a] += p0 + p1
        return tuple(s)

    def __call__(self, x):
        return self._pad(np.asarray(x), self.padding, self.data_format)


class OZeroPadding3D(OPeriodicPadding3D):
    _pad = staticmethod(ops.zero_pad3d)


class OConvLSTM2D(OLayer):
    """keras ConvLSTM2D as built at examples/train.py:150-157 (channels_first, (T, C, H, W) per sample); numpy only."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', data_format=None, dilation_rate=(1, 1),
                 activation='tanh', recurrent_activation='hard_sigmoid', use_bias=True, unit_forget_bias=True,
                 return_sequences=False, **kwargs):
        if (data_format or 'channels_last') != 'channels_first' or ops.normalize_pair(strides) != (1, 1):
            raise ValueError('oracle ConvLSTM2D restates the channels_first, stride-1 layer of the example nets')
        self.filters = int(filters)
        self.kernel_size = ops.normalize_pair(kernel_size, 'kernel_size')
        self.dilation_rate = ops.normalize_pair(dilation_rate, 'dilation_rate')
        self.padding = padding
        self.activation, self.recurrent_activation = activation, recurrent_activation
        self.use_bias, self.unit_forget_bias = use_bias, unit_forget_bias
        self.return_sequences = return_sequences
        self.kernel = self.recurrent_kernel = self.bias = None

    def build(self, s):
        kh, kw = self.kernel_size
        F = self.filters
        self.kernel = np.zeros((kh, kw, s[1], 4 * F), np.float32)
        self.recurrent_kernel = np.zeros((kh, kw, F, 4 * F), np.float32)
        if self.use_bias:
            self.bias = np.zeros((4 * F,), np.float32)
            if self.unit_forget_bias:
                self.bias[F:2 * F] = 1.0

    def output_shape(self, s):
        kh, kw = self.kernel_size
        dh, dw = self.dilation_rate
        H, W = (s[2], s[3]) if self.padding == 'same' else (s[2] - dh * (kh - 1), s[3] - dw * (kw - 1))
        return (s[0], self.filters, H, W) if self.return_sequences else (self.filters, H, W)

    @property
    def weights(self):
        return [self.kernel, self.recurrent_kernel] + ([self.bias] if self.use_bias else [])

    def set_weights(self, ws):
        self.kernel, self.recurrent_kernel = np.asarray(ws[0]), np.asarray(ws[1])
        if self.use_bias:
            self.bias = np.asarray(ws[2])

    def randomize(self, rng, bias_scale=0.0):
        kh, kw, cin, f4 = self.kernel.shape
        self.kernel = ops.glorot_uniform(rng, kh, kw, cin, f4)
        self.recurrent_kernel = ops.glorot_uniform(rng, kh, kw, self.filters, f4)
        if self.use_bias and bias_scale:
            self.bias = (self.bias + bias_scale * rng.standard_normal(self.bias.shape)).astype(np.float32)

    def __call__(self, x):
        x = np.asarray(x)
        cast = lambda w: None if w is None else np.asarray(w, x.dtype)
        return ops.conv_lstm2d(x, cast(self.kernel), cast(self.recurrent_kernel), cast(self.bias), self.dilation_rate,
                               self.padding, self.activation, self.recurrent_activation, self.return_sequences)


class OReshape(OLayer):
    def __init__(self, target_shape, **kwargs):
        self.target_shape = tuple(target_shape)

    def output_shape(self, s):
        return self.target_shape

    def __call__(self, x):
        return x.reshape((x.shape[0],) + self.target_shape)


class OSlice(OLayer):
    """slice_layer(start, end, step, axis) -- DLWP/custom.py:675-692."""

    def __init__(self, start, end, step=None, axis=1):
        if axis < 0:
            raise ValueError("'slice_layer' can only work on a specified axis > 0")
        self.start, self.end, self.step, self.axis = start, end, step, axis

    def output_shape(self, s):
        s = list(s)
        s[self.axis - 1] = len(range(*slice(self.start, self.end, self.step).indices(s[self.axis - 1])))
        return tuple(s)

    def __call__(self, x):
        idx = [slice(None)] * self.axis + [slice(self.start, self.end, self.step)]
        return x[tuple(idx)]


def concatenate(xs, axis=1):
    if _is_torch(xs[0]):
        return torch.cat(list(xs), dim=axis)
    return np.concatenate(list(xs), axis=axis)


LAYER_REGISTRY = {
    'PeriodicPadding2D': OPeriodicPadding2D,
    'ZeroPadding2D': OZeroPadding2D,
    'FillPadding2D': OFillPadding2D,
    'TFPadding2D': OTFPadding2D,
    'Conv2D': OConv2D,
    'RowConnected2D': ORowConnected2D,
    'MaxPooling2D': OMaxPooling2D,
    'UpSampling2D': OUpSampling2D,
    'Reshape': OReshape,
    'PeriodicPadding3D': OPeriodicPadding3D,
    'ZeroPadding3D': OZeroPadding3D,
    'ConvLSTM2D': OConvLSTM2D,
}


def init_weights(conv_layers, seed=1, bias_scale=0.0):
    """
    BASELINE.md section 3: kernels = Keras default glorot_uniform restated with RandomState(seed), drawn per layer in
    (kh,kw,Cin,Cout) order; biases zero (or N(0, bias_scale) for tests that want to exercise the bias path).
    """
    rng = np.random.RandomState(seed)
    for layer in conv_layers:
        shp = layer.kernel.shape
        if len(shp) == 4:
            layer.kernel = ops.glorot_uniform(rng, *shp)
        else:  # row-connected: one glorot draw per row
            layer.kernel = np.stack([ops.glorot_uniform(rng, *shp[1:]) for _ in range(shp[0])], axis=0)
        if layer.use_bias:
            if bias_scale:
                layer.bias = (bias_scale * rng.standard_normal(layer.bias.shape)).astype(np.float32)
            else:
                layer.bias = np.zeros(layer.bias.shape, np.float32)
        layer._tw = None


class OSequential(object):
    """The layer-tuple interpreter: DLWP/model/models.py:74-103."""

    def __init__(self, layers, input_shape=None):
        if type(layers) not in [list, tuple]:
            raise TypeError("'layers' argument must be a tuple")
        self.layers = []
        for l, layer in enumerate(layers):
            if type(layer) not in [list, tuple]:
                raise TypeError("each element of 'layers' must be a tuple")
            if len(layer) != 3:
                raise ValueError("each layer must be specified by three elements (name, args, kwargs)")
            name, args, kwargs = layer
            args = () if args is None else args
            kwargs = {} if kwargs is None else dict(kwargs)
            if type(args) is not tuple:
                raise TypeError("the 'args' element of layer %d must be a tuple" % l)
            if type(kwargs) is not dict:
                raise TypeError("the 'kwargs' element of layer %d must be a dict" % l)
            if 'input_shape' in kwargs:
                shp = kwargs.pop('input_shape')
                if l == 0 and input_shape is None:
                    input_shape = shp
            self.layers.append(LAYER_REGISTRY[name](*args, **kwargs))
        if input_shape is None:
            raise ValueError('input_shape is required')
        self.input_shape = tuple(input_shape)
        s = self.input_shape
        for layer in self.layers:
            layer.build(s)
            s = layer.output_shape(s)
        self.output_shape = s
        self.n_outputs = 1

    @property
    def conv_layers(self):
        return [l for l in self.layers if isinstance(l, OConv2D)]

    @property
    def weight_layers(self):
        return [l for l in self.layers if isinstance(l, (OConv2D, OConvLSTM2D))]

    def get_weights(self):
        return [w for l in self.weight_layers for w in l.weights]

    def set_weights(self, ws):
        ws = list(ws)
        for l in self.weight_layers:
            n = len(l.weights)
            l.set_weights(ws[:n])
            ws = ws[n:]

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class OFunctionalNet(object):
    """
    examples/train_functional.py:154-285, non-recurrent: ``basic_model`` (222-245) or ``skip_model`` (248-275) with
    shared layers, applied ``integration_steps`` times (278-281).  ``forward`` returns the list of unrolled outputs
    (a bare array for integration_steps == 1, like keras Model.predict).
    """

    def __init__(self, cs, cso=None, skip_connections=True, integration_steps=1, latitude_dependent=False):
        cso = cs if cso is None else cso
        self.cs, self.cso = tuple(cs), tuple(cso)
        self.skip = skip_connections
        self.n_outputs = int(integration_steps)
        cf = 'channels_first'
        self.pp2 = OPeriodicPadding2D((0, 2), cf)
        self.zp2 = OZeroPadding2D((2, 0), cf)
        self.pp1 = OPeriodicPadding2D((0, 1), cf)
        self.zp1 = OZeroPadding2D((1, 0), cf)
        self.pool = OMaxPooling2D(2, data_format=cf)
        self.up = OUpSampling2D(2, data_format=cf)
        self.c1 = OConv2D(32, 3, dilation_rate=2, activation='tanh', data_format=cf)
        self.c2 = OConv2D(64, 3, dilation_rate=1, activation='tanh', data_format=cf)
        self.c3 = OConv2D(128, 3, dilation_rate=1, activation='tanh', data_format=cf)
        self.c4 = OConv2D(32 if self.skip else 64, 3, dilation_rate=1, activation='tanh', data_format=cf)
        self.c5 = OConv2D(16 if self.skip else 32, 3, dilation_rate=2, activation='tanh', data_format=cf)
        last = ORowConnected2D if latitude_dependent else OConv2D
        self.c6 = last(self.cso[0], 5, activation='linear', data_format=cf)
        self.s11, self.s12 = OSlice(0, 16, axis=1), OSlice(16, 32, axis=1)
        self.s21, self.s22 = OSlice(0, 32, axis=1), OSlice(32, 64, axis=1)
        self.conv_layers = [self.c1, self.c2, self.c3, self.c4, self.c5, self.c6]
        C, H, W = self.cs
        ins = [(C, H + 4, W + 4), (16 if self.skip else 32, H // 2 + 2, W // 2 + 2),
               (32 if self.skip else 64, H // 4 + 2, W // 4 + 2), (128, H // 2 + 2, W // 2 + 2),
               (64, H + 4, W + 4), (32, H + 4, W + 4)]
        for layer, s in zip(self.conv_layers, ins):
            layer.build(s)
        self.input_shape = self.cs
        self.output_shape = self.cso

    def get_weights(self):
        return [w for l in self.conv_layers for w in l.weights]

    def set_weights(self, ws):
        ws = list(ws)
        for l in self.conv_layers:
            n = len(l.weights)
            l.set_weights(ws[:n])
            ws = ws[n:]

    def _basic(self, x):
        x = self.c1(self.pp2(self.zp2(x)))
        x = self.pool(x)
        x = self.c2(self.pp1(self.zp1(x)))
        x = self.pool(x)
        x = self.c3(self.pp1(self.zp1(x)))
        x = self.up(x)
        x = self.c4(self.pp1(self.zp1(x)))
        x = self.up(x)
        x = self.c5(self.pp2(self.zp2(x)))
        x = self.c6(self.pp2(self.zp2(x)))
        return x

    def _skip(self, x):
        x = self.c1(self.pp2(self.zp2(x)))
        x, x1 = self.s11(x), self.s12(x)
        x = self.pool(x)
        x = self.c2(self.pp1(self.zp1(x)))
        x, x2 = self.s21(x), self.s22(x)
        x = self.pool(x)
        x = self.c3(self.pp1(self.zp1(x)))
        x = self.up(x)
        x = self.c4(self.pp1(self.zp1(x)))
        x = concatenate([x, x2], axis=1)
        x = self.up(x)
        x = self.c5(self.pp2(self.zp2(x)))
        x = concatenate([x, x1], axis=1)
        x = self.c6(self.pp2(self.zp2(x)))
        return x

    def forward(self, x):
        f = self._skip if self.skip else self._basic
        outs = [f(x)]
        for _ in range(1, self.n_outputs):
            outs.append(f(outs[-1]))
        return outs[0] if self.n_outputs == 1 else outs


def net_a_layers(cs=(6, 91, 180)):
    """
    "Net A" of SURVEY.md section 8 / BASELINE.json configs[0-1]: first and last conv blocks of examples/train.py:159-169,
    211-219 (channels_first, time_dim=1).
    """
    cf = 'channels_first'
    return (
        ('PeriodicPadding2D', ((0, 2),), {'data_format': cf, 'input_shape': tuple(cs)}),
        ('ZeroPadding2D', ((2, 0),), {'data_format': cf}),
        ('Conv2D', (32, 3), {'dilation_rate': 2, 'padding': 'valid', 'activation': 'tanh', 'data_format': cf}),
        ('PeriodicPadding2D', ((0, 2),), {'data_format': cf}),
        ('ZeroPadding2D', ((2, 0),), {'data_format': cf}),
        ('Conv2D', (cs[0], 5), {'padding': 'valid', 'activation': 'linear', 'data_format': cf}),
    )
