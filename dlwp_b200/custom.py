"""
Mirror of `DLWP.custom` (reference DLWP/custom.py): the custom layer, loss and callback NAMES that the example scripts
import and that `DLWPNeuralNet.build_model` resolves by string (DLWP/model/models.py:97-103).

Layers only describe geometry; their arithmetic is fused into the convolution kernels of libdlwp_b200.so by
dlwp_b200.engine (PeriodicPadding2D never materialises a padded copy -- compare custom.py:202-204).
"""

import numpy as np

from .keras.callbacks import Callback, EarlyStopping
from .keras.layers import (ChannelSlice, Conv2D, Layer, ZeroPadding2D, ZeroPadding3D, _glorot_uniform, _norm_tuple)
from .keras.losses import mean_absolute_error, mean_squared_error


# ==================================================================================================================== #
# Layers
# ==================================================================================================================== #

class PeriodicPadding2D(ZeroPadding2D):
    """
    Periodic (torus) padding for 2-D inputs: DLWP/custom.py:139-214.  Same constructor as keras ZeroPadding2D
    (`padding` int / (sym_h, sym_w) / ((top, bottom), (left, right)); `data_format`).  W is wrapped first, then H of the
    W-padded tensor, so corners wrap in both dimensions (custom.py:201-204).
    """
    pad_mode = 'periodic'

    def __init__(self, padding=(1, 1), data_format=None, **kwargs):
        super(PeriodicPadding2D, self).__init__(padding=padding, data_format=data_format, **kwargs)


class PeriodicPadding3D(ZeroPadding3D):
    """DLWP/custom.py:217-306: periodic padding of the three trailing axes of a 5-D input; on the recurrent nets'
    (batch, time, channels, lat, lon) tensors `padding=(0, 0, p)` wraps longitude (examples/train.py:144-148).  Lowered
    as a pending 2-D padding of every (time, channel) plane, fused into the ConvLSTM2D input convolution."""
    pad_mode = 'periodic'

    def __init__(self, padding=(1, 1, 1), data_format=None, **kwargs):
        super(PeriodicPadding3D, self).__init__(padding=padding, data_format=data_format, **kwargs)


class FillPadding2D(ZeroPadding2D):
    """DLWP/custom.py:309-402: replicates the border rows, then the border columns of the row-padded tensor (corners =
    corner value).  Runs as a stand-alone pad op (DLWP_PAD_EDGE); not used by the example nets."""
    pad_mode = 'fill'


class FillPadding3D(ZeroPadding3D):
    pad_mode = 'fill'


class TFPadding2D(ZeroPadding2D):
    """DLWP/custom.py:527-599: tf.pad with mode CONSTANT (0 only: equals ZeroPadding2D), REFLECT or SYMMETRIC; the
    mirrored modes run as stand-alone pad ops (DLWP_PAD_REFLECT / DLWP_PAD_SYMMETRIC)."""

    def __init__(self, padding=(1, 1), data_format=None, mode='CONSTANT', constant_values=0, **kwargs):
        super(TFPadding2D, self).__init__(padding=padding, data_format=data_format, **kwargs)
        self.mode = mode
        self.constant_values = constant_values
        if mode.upper() == 'CONSTANT':
            self.pad_mode = 'zero' if constant_values == 0 else 'constant(%r)' % (constant_values,)
        else:
            self.pad_mode = mode.lower()

    def get_config(self):
        c = super(TFPadding2D, self).get_config()
        c.update({'mode': self.mode, 'constant_values': self.constant_values})
        return c


class TFPadding3D(ZeroPadding3D):
    def __init__(self, padding=(1, 1, 1), data_format=None, mode='CONSTANT', constant_values=0, **kwargs):
        super(TFPadding3D, self).__init__(padding=padding, data_format=data_format, **kwargs)
        self.mode = mode
        self.constant_values = constant_values


def slice_layer(start, end, step=None, axis=1):
    """
    DLWP/custom.py:675-692: a Lambda layer slicing `x[:, start:end:step]` along `axis`.  Returns a `ChannelSlice`
    (a Lambda subclass) so the GPU plan can treat it as a zero-copy channel window.
    """
    if axis < 0:
        raise ValueError("'slice_layer' can only work on a specified axis > 0")
    return ChannelSlice(start, end, step, axis)


class RowConnected2D(Conv2D):
    """
    Row-connected layer (weights shared only along rows): DLWP/custom.py:695-837.  kernel (H_out, kh, kw, Cin, Cout),
    bias (H_out, 1, Cout); only padding='valid' (custom.py:734) and strides (1, 1) (custom.py:881 mixes the two
    strides) are defined.
    """

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', data_format=None, activation=None,
                 use_bias=True, **kwargs):
        for k in ('kernel_initializer', 'bias_initializer', 'kernel_regularizer', 'bias_regularizer',
                  'activity_regularizer', 'kernel_constraint', 'bias_constraint'):
            kwargs.pop(k, None)
        if str(padding).lower() != 'valid':
            raise ValueError('Invalid border mode for RowConnected2D (only "valid" is supported): ' + str(padding))
        super(RowConnected2D, self).__init__(filters, kernel_size, strides=strides, padding='valid',
                                             data_format=data_format, activation=activation, use_bias=use_bias,
                                             **kwargs)

    def build(self, s):
        hi, wi = (2, 3) if self.data_format == 'channels_first' else (1, 2)
        if s[hi] is None or s[wi] is None:
            raise ValueError('The spatial dimensions of the inputs to  a LocallyConnected2D layer should be '
                             'fully-defined, but layer received the inputs shape ' + str(s))
        cin = s[self._channel_axis()]
        kh, kw = self.kernel_size
        self.output_row = (s[hi] - kh) // self.strides[0] + 1
        self.output_col = (s[wi] - kw) // self.strides[1] + 1
        rng = np.random.RandomState(np.random.randint(0, 2 ** 31 - 1))
        kernel = np.stack([_glorot_uniform(rng, (kh, kw, cin, self.filters), kh * kw * cin, kh * kw * self.filters)
                           for _ in range(self.output_row)], axis=0)
        self._weights = [kernel] + ([np.zeros((self.output_row, 1, self.filters), np.float32)]
                                    if self.use_bias else [])
        self.built = True


# ==================================================================================================================== #
# Losses (host-side definitions; the training path maps the names to device code)
# ==================================================================================================================== #

class _LatLoss(object):
    """The `lat_loss` closure of DLWP/custom.py:986-990 as a picklable callable (util.save_model pickles the compile
    arguments; a nested function cannot be pickled)."""
    __name__ = 'lat_loss'

    def __init__(self, base_loss, weights):
        self.base_loss = base_loss
        self.weights = np.asarray(weights, np.float32)

    def __call__(self, y_true, y_pred):
        return self.base_loss(y_true * self.weights, y_pred * self.weights)


def latitude_weighted_loss(loss_function=mean_squared_error, lats=None, output_shape=(), axis=-2, weighting='cosine'):
    """
    DLWP/custom.py:956-991: multiply y_true and y_pred by a function of latitude before `loss_function`.
    weights = cos(lat) [+ 0.5 sin^2(2 lat) for 'midlatitude' (custom.py:976-978)], broadcast over the trailing axes.
    """
    if weighting not in ['cosine', 'midlatitude']:
        raise ValueError("'weighting' must be one of 'cosine' or 'midlatitude'")
    if lats is not None:
        lat = np.asarray(lats, dtype=np.float32)
        weights = np.cos(lat * np.pi / 180.)
        if weighting == 'midlatitude':
            weights = weights + 0.5 * np.power(np.sin(lat * 2 * np.pi / 180.), 2.)
        weight_shape = tuple(output_shape)[axis:]
        for d in weight_shape[1:]:
            weights = np.repeat(np.expand_dims(weights, axis=-1), d, axis=-1)
    else:
        weights = np.ones(tuple(output_shape), np.float32)
    return _LatLoss(loss_function, weights)


def _acc_terms(y_true, y_pred, mean, regularize_mean):
    """(a, m) of DLWP/custom.py:1060-1075: the correlation of the anomalies (climatology `mean` subtracted, None = 0) and the
    regularizer, which the reference computes on the RAW tensors."""
    if regularize_mean is not None:
        assert regularize_mean in ['global', 'spatial', 'mse', 'mae']
    pa, ta = (y_pred, y_true) if mean is None else (y_pred - mean, y_true - mean)
    a = np.mean(pa * ta) / np.sqrt(np.mean(np.square(pa)) * np.mean(np.square(ta)))
    m = None
    if regularize_mean == 'global':
        m = np.abs((np.mean(y_true) - np.mean(y_pred)) / np.mean(y_true))
    elif regularize_mean == 'spatial':
        m = np.mean(np.abs((np.mean(y_true, axis=(-2, -1)) - np.mean(y_pred, axis=(-2, -1)))
                           / np.mean(y_true, axis=(-2, -1))))
    elif regularize_mean == 'mse':
        m = mean_squared_error(y_true, y_pred)
    elif regularize_mean == 'mae':
        m = mean_absolute_error(y_true, y_pred)
    return a, m


def anomaly_correlation(y_true, y_pred, mean=0., regularize_mean='mse', reverse=True):
    """DLWP/custom.py:994-1033 on numpy arrays (climatological mean assumed 0: `mean` is ignored, as in the reference)."""
    a, m = _acc_terms(y_true, y_pred, None, regularize_mean)
    if reverse:
        return m - a if regularize_mean is not None else -a
    return a - m if regularize_mean else a


class _AccLoss(object):
    """The `acc_loss` closure of DLWP/custom.py:1057-1086 as a picklable callable."""
    __name__ = 'acc_loss'

    def __init__(self, mean, regularize_mean, reverse):
        self.mean, self.regularize_mean, self.reverse = mean, regularize_mean, reverse

    def __call__(self, y_true, y_pred):
        a, m = _acc_terms(y_true, y_pred, self.mean, self.regularize_mean)
        if self.reverse:
            return m - a if self.regularize_mean is not None else -a
        return a - m if self.regularize_mean else a


def anomaly_correlation_loss(mean=None, regularize_mean='mse', reverse=True):
    """DLWP/custom.py:1036-1088."""
    if mean is not None:
        assert len(mean.shape) > 1
        assert mean.shape[0] == 1
    if regularize_mean is not None:
        assert regularize_mean in ['global', 'spatial', 'mse', 'mae']
        reverse = True
    return _AccLoss(mean, regularize_mean, reverse)


# Defined at import time so that models saved with these losses can be re-loaded by name (custom.py:1092-1093)
lat_loss = latitude_weighted_loss()
acc_loss = anomaly_correlation_loss()


# ==================================================================================================================== #
# Callbacks (host-side training bookkeeping: custom.py:32-136)
# ==================================================================================================================== #

class AdamLearningRateTracker(Callback):
    def on_epoch_end(self, epoch, logs=None, beta_1=0.9, beta_2=0.999):
        opt = self.model.optimizer
        t = opt.iterations + 1.
        new_lr = opt.lr * (1. / (1. + opt.decay * opt.iterations))
        lr_t = new_lr * (np.sqrt(1. - np.power(beta_2, t)) / (1. - np.power(beta_1, t)))
        print(' - LR: {:.6f}'.format(lr_t))


class SGDLearningRateTracker(Callback):
    def on_epoch_end(self, epoch, logs=None):
        opt = self.model.optimizer
        print(' - LR: {:.6f}'.format(opt.lr * (1. / (1. + opt.decay * opt.iterations))))


class BatchHistory(Callback):
    def on_train_begin(self, logs=None):
        self.history = []
        self.epoch = 0

    def on_epoch_begin(self, epoch, logs=None):
        self.history.append({})

    def on_epoch_end(self, epoch, logs=None):
        self.epoch += 1

    def on_batch_end(self, batch, logs=None):
        for k, v in (logs or {}).items():
            self.history[self.epoch].setdefault(k, []).append(v)


class RunHistory(Callback):
    """History that also logs to an (Azure) run object: custom.py:71-91."""

    def __init__(self, run):
        super(RunHistory, self).__init__()
        self.epoch = []
        self.history = {}
        self.run = run

    def on_train_begin(self, logs=None):
        self.epoch = []
        self.history = {}

    def on_epoch_end(self, epoch, logs=None):
        self.epoch.append(epoch)
        for k, v in (logs or {}).items():
            self.history.setdefault(k, []).append(v)
            self.run.log(k, v)


class RNNResetStates(Callback):
    def on_epoch_begin(self, epoch, logs=None):
        self.model.reset_states()


class EarlyStoppingMin(EarlyStopping):
    """EarlyStopping that trains for at least `min_epochs` epochs first: custom.py:99-136."""

    def __init__(self, min_epochs=0, **kwargs):
        super(EarlyStoppingMin, self).__init__(**kwargs)
        if not isinstance(min_epochs, int) or min_epochs < 0:
            raise ValueError('min_epochs must be an integer >= 0')
        self.min_epochs = min_epochs

    def on_epoch_end(self, epoch, logs=None):
        if epoch < self.min_epochs:
            return
        super(EarlyStoppingMin, self).on_epoch_end(epoch, logs)


# ==================================================================================================================== #
# PyTorch classes (custom.py:1098-1107)
# ==================================================================================================================== #

class TorchReshape(object):
    """`x.view(*shape)` as a layer of a DLWPTorchNN (the torch twin resolves unknown layer names here)."""

    def __init__(self, shape):
        if not isinstance(shape, tuple):
            raise ValueError("'shape' must be a tuple of integers")
        self.shape = shape

    def __call__(self, x):
        return x.view(*self.shape)
