"""
Mirror of `DLWP.model.models` (reference DLWP/model/models.py): `DLWPNeuralNet` and `DLWPFunctional` with the same
constructor arguments, attributes, methods, exceptions and numpy-in / numpy-out contracts -- so the reference's example
and validation scripts are drop-in -- but with `predict_timeseries` executed as ONE device-resident rollout
(libdlwp_b200: dlwp_rollout_host) instead of a Python loop of Keras predict calls with three host copies of the state
per step (models.py:271, 292, 293).
"""

import numpy as np

from .. import keras, util
from .generators import DataGenerator, SeriesDataGenerator, SmartDataGenerator


def _series_reshape(series, steps, n, time_dim, feature_shape, keep_time_dim, first_slice_only=False):
    """
    The output rule shared by both drivers (models.py:294-300, :448-451): (steps, N, T*C, ...) -> (steps, N, T, C, ...)
    and, unless keep_time_dim, fold T into the leading forecast-step axis (or keep only slice 0 for step_sequence).
    """
    series = series.reshape((steps, n, time_dim, -1) + tuple(feature_shape[1:]))
    if not keep_time_dim:
        if first_slice_only:
            series = series[:, :, 0]
        else:
            series = series.transpose((0, 2, 1) + tuple(range(3, 3 + len(feature_shape))))
            series = series.reshape((steps * time_dim, n, -1) + tuple(feature_shape[1:]))
    return series


class DLWPNeuralNet(object):
    """
    DLWP model class around a Sequential network built from `(layer_name, args, kwargs)` tuples
    (reference DLWP/model/models.py:21-317).
    """

    def __init__(self, is_convolutional=True, is_recurrent=False, time_dim=1, scaler_type='StandardScaler',
                 scale_targets=True, apply_same_y_scaling=True, impute_missing=False):
        self.is_convolutional = is_convolutional
        self.is_recurrent = is_recurrent
        if int(time_dim) < 1:
            raise ValueError("'time_dim' must be >= 1")
        self.time_dim = time_dim
        self.scaler_type = scaler_type
        self.scale_targets = scale_targets
        self.apply_same_y_scaling = apply_same_y_scaling
        self.scaler = None
        self.scaler_y = None
        self.impute = impute_missing
        self.imputer = None
        self.imputer_y = None
        self.base_model = None
        self.model = None
        self.gpus = 1
        self._is_init_fit = scaler_type is None

    # -- construction (models.py:63-112) -------------------------------------------------------------------------
    def build_model(self, layers=(), gpus=1, **compile_kwargs):
        """
        Build the Sequential network.  Each element of `layers` is `(layer_name, args_tuple_or_None,
        kwargs_dict_or_None)`; names resolve in `keras.layers` first and `DLWP.custom` second (models.py:97-103).
        `gpus` > 1 wraps the model in `keras.utils.multi_gpu_model` like models.py:104-109: predict / predict_timeseries
        split every batch over that many devices of this process.
        """
        if type(gpus) is not int:
            raise TypeError("'gpus' argument must be an int")
        if type(layers) not in [list, tuple]:
            raise TypeError("'layers' argument must be a tuple")
        specs = []
        for l, layer in enumerate(layers):
            if type(layer) not in [list, tuple]:
                raise TypeError("each element of 'layers' must be a tuple")
            if len(layer) != 3:
                raise ValueError("each layer must be specified by three elements (name, args, kwargs)")
            name, args, kwargs = layer
            args = () if args is None else args
            kwargs = {} if kwargs is None else kwargs
            if type(args) is not tuple:
                raise TypeError("the 'args' element of layer %d must be a tuple" % l)
            if type(kwargs) is not dict:
                raise TypeError("the 'kwargs' element of layer %d must be a dict" % l)
            specs.append((name, args, kwargs))
        util.make_keras_picklable()
        self.base_model = keras.models.Sequential()
        for name, args, kwargs in specs:
            try:
                layer_class = util.get_from_class('keras.layers', name)
            except (ImportError, AttributeError):
                layer_class = util.get_from_class('DLWP.custom', name)
            self.base_model.add(layer_class(*args, **kwargs))
        self.model = keras.utils.multi_gpu_model(self.base_model, gpus=gpus) if gpus > 1 else self.base_model
        self.gpus = gpus
        self.model.compile(**compile_kwargs)

    # -- sklearn pre-processing (models.py:114-194) ----------------------------------------------------------------
    @staticmethod
    def _reshape(a, ret=False):
        shp = a.shape
        a = a.reshape((shp[0], -1))
        return (a, shp) if ret else a

    def scaler_fit(self, X, y, **kwargs):
        if self.scaler_type is None:
            return
        cls = util.get_from_class('sklearn.preprocessing', self.scaler_type)
        self.scaler = cls(**kwargs)
        self.scaler_y = cls(**kwargs)
        self.scaler.fit(self._reshape(X))
        if self.scale_targets:
            if self.apply_same_y_scaling:
                self.scaler_y = self.scaler
            else:
                self.scaler_y.fit(self._reshape(y))

    def scaler_transform(self, X, y=None):
        if self.scaler_type is None:
            return (X, y) if y is not None else X
        X2, xs = self._reshape(X, ret=True)
        Xt = self.scaler.transform(X2).reshape(xs)
        if y is None:
            return Xt
        if self.scale_targets:
            y2, ys = self._reshape(y, ret=True)
            return Xt, self.scaler_y.transform(y2).reshape(ys)
        return Xt, y

    def imputer_fit(self, X, y):
        try:
            cls = util.get_from_class('sklearn.preprocessing', 'Imputer')
            make = lambda: cls(missing_values=np.nan, strategy="mean", axis=0, copy=False)
        except (ImportError, AttributeError):  # sklearn >= 0.22 renamed it
            cls = util.get_from_class('sklearn.impute', 'SimpleImputer')
            make = lambda: cls(missing_values=np.nan, strategy="mean", copy=False)
        self.imputer = make()
        self.imputer_y = make()
        self.imputer.fit(self._reshape(X))
        if self.apply_same_y_scaling:
            self.imputer_y = self.imputer
        else:
            self.imputer_y.fit(self._reshape(y))

    def imputer_transform(self, X, y=None):
        X2, xs = self._reshape(X, ret=True)
        Xt = self.imputer.transform(X2).reshape(xs)
        if y is None:
            return Xt
        y2, ys = self._reshape(y, ret=True)
        return Xt, self.imputer_y.transform(y2).reshape(ys)

    def init_fit(self, predictors, targets, scaler_kwargs=None):
        scaler_kwargs = scaler_kwargs or {}
        if self.impute:
            self.imputer_fit(predictors, targets)
            predictors, targets = self.imputer_transform(predictors, y=targets)
        self.scaler_fit(predictors, targets, **scaler_kwargs)
        self._is_init_fit = True

    # -- training (models.py:188-228) ------------------------------------------------------------------------------
    def fit(self, predictors, targets, initialize=True, **kwargs):
        if initialize:
            self.init_fit(predictors, targets)
        elif not self._is_init_fit:
            raise AttributeError('DLWPNeuralNet has not been initialized for fitting with init_fit()')
        if self.impute:
            predictors, targets = self.imputer_transform(predictors, y=targets)
        ps, ts = self.scaler_transform(predictors, targets)
        if kwargs.get('validation_data') is not None:
            vp, vt = kwargs['validation_data']
            if self.impute:
                vp, vt = self.imputer_transform(vp, vt)
            kwargs['validation_data'] = self.scaler_transform(vp, vt)
        self.model.fit(ps, ts, **kwargs)

    def fit_generator(self, generator, **kwargs):
        if isinstance(generator, (DataGenerator, SmartDataGenerator, SeriesDataGenerator)):
            if not self._is_init_fit:
                raise AttributeError('DLWPNeuralNet has not been initialized for fitting with init_fit()')
        self.model.fit_generator(generator, **kwargs)

    # -- inference (models.py:230-301) -----------------------------------------------------------------------------
    def predict(self, predictors, **kwargs):
        """models.py:230-245: optional impute + scale, model.predict, optional inverse scale of the targets."""
        if self.impute:
            predictors = self.imputer_transform(predictors)
        predicted = self.model.predict(self.scaler_transform(predictors), **kwargs)
        if self.scale_targets and self.scaler_type is not None:
            return self.scaler_y.inverse_transform(predicted)
        return predicted

    def _device_rollout_ok(self, predictors, step_sequence):
        return (not self.impute and self.scaler_type is None and
                hasattr(self.model, 'engine') and predictors.ndim == (5 if self.is_recurrent else 4) and
                (not step_sequence or len(self.model.outputs) == 1) and
                self.model.engine(predictors.shape[0]).can_rollout())

    def predict_timeseries(self, predictors, time_steps, step_sequence=False, keep_time_dim=False, **kwargs):
        """
        models.py:247-301, same arguments and return value: ceil(time_steps / time_dim) model applications (NOT
        truncated to time_steps), each fed the previous output; returns float32 (steps*time_dim, N, C, H, W), or
        (steps, N, time_dim, C, H, W) with keep_time_dim.  `step_sequence=True` advances one time slice per application
        (models.py:280-290).  `verbose` in kwargs prints progress like the reference when the host loop is used.

        The common case (no scaler/imputer, not recurrent, not step_sequence) runs as one device-resident rollout; the
        other cases keep the reference's host loop around GPU `predict` calls.
        """
        time_steps = int(time_steps)
        if time_steps < 1:
            raise ValueError("time_steps must be an int > 0")
        predictors = np.asarray(predictors)
        if not step_sequence:
            time_steps = int(np.ceil(1. * time_steps / self.time_dim))
        sample_dim = predictors.shape[0]
        feature_shape = predictors.shape[2:] if self.is_recurrent else predictors.shape[1:]

        if self._device_rollout_ok(predictors, step_sequence):
            if step_sequence:     # models.py:280-290 as one device-resident loop (one time slice per application)
                series = self.model.rollout_step_sequence_host(predictors, time_steps, self.time_dim)
            else:
                series = self.model.rollout_host(predictors, time_steps)
        else:
            series = np.full((time_steps,) + predictors.shape, np.nan, dtype=np.float32)
            p = predictors.copy()
            for t in range(time_steps):
                if kwargs.get('verbose', 0) > 0:
                    print('Time step %d/%d' % (t + 1, time_steps))
                if step_sequence:
                    pr = self.predict(p, **kwargs)
                    pr_shape = pr.shape
                    if not self.is_recurrent:
                        pr = pr.reshape((sample_dim, self.time_dim, -1) + feature_shape[1:])
                        p = p.reshape((sample_dim, self.time_dim, -1) + feature_shape[1:])
                    p = np.concatenate((p[:, 1:], pr[:, [0]]), axis=1)
                    if not self.is_recurrent:
                        p = p.reshape(predictors.shape)
                        pr = pr.reshape(pr_shape)
                    series[t, ...] = pr
                else:
                    p = self.predict(p, **kwargs)
                    series[t, ...] = p
        return _series_reshape(series, time_steps, sample_dim, self.time_dim, feature_shape, keep_time_dim,
                               first_slice_only=step_sequence)

    def evaluate(self, predictors, targets, **kwargs):
        if self.impute:
            predictors, targets = self.imputer_transform(predictors, targets)
        ps, ts = self.scaler_transform(predictors, targets)
        return self.model.evaluate(ps, ts, **kwargs)


class DLWPFunctional(object):
    """
    DLWP model class around a user-built functional `keras.models.Model`, possibly with several outputs = several
    unrolled applications of a shared-weight net (reference DLWP/model/models.py:319-464).  No scaling / imputing.
    """

    def __init__(self, is_convolutional=True, is_recurrent=False, time_dim=1):
        self.is_convolutional = is_convolutional
        self.is_recurrent = is_recurrent
        if int(time_dim) < 1:
            raise ValueError("'time_dim' must be >= 1")
        self.time_dim = time_dim
        self.scaler = None
        self.scaler_y = None
        self.impute = False
        self.imputer = None
        self.imputer_y = None
        self._n_steps = 1
        self.base_model = None
        self.model = None
        self.gpus = 1

    def build_model(self, model, gpus=1, **compile_kwargs):
        """models.py:349-373: adopt a functional model; `_n_steps = len(model.outputs)` (models.py:364)."""
        if type(gpus) is not int:
            raise TypeError("'gpus' argument must be an int")
        util.make_keras_picklable()
        self.base_model = model
        self._n_steps = len(model.outputs)
        self.model = keras.utils.multi_gpu_model(self.base_model, gpus=gpus) if gpus > 1 else self.base_model
        self.gpus = gpus
        self.model.compile(**compile_kwargs)

    def scaler_transform(self, X, y=None):
        return (X, y) if y is not None else X

    def fit(self, predictors, targets, **kwargs):
        self.model.fit(predictors, targets, **kwargs)

    def fit_generator(self, generator, **kwargs):
        self.model.fit_generator(generator, **kwargs)

    def predict(self, predictors, **kwargs):
        return self.model.predict(predictors, **kwargs)

    def predict_timeseries(self, predictors, time_steps, keep_time_dim=False, **kwargs):
        """
        models.py:414-452: ceil(time_steps / _n_steps / time_dim) model applications, each producing _n_steps outputs of
        which the LAST is fed back (models.py:443-446) and all are stored (447).  Returns float32
        (out_steps*time_dim, N, C, H, W), or (out_steps, N, time_dim, C, H, W) with keep_time_dim.
        """
        time_steps = int(time_steps)
        if time_steps < 1:
            raise ValueError("time_steps must be an int > 0")
        predictors = np.asarray(predictors)
        steps = int(np.ceil(time_steps / self._n_steps / self.time_dim))
        out_steps = steps * self._n_steps
        sample_dim = predictors.shape[0]
        feature_shape = predictors.shape[2:] if self.is_recurrent else predictors.shape[1:]
        eng = self.model.engine(sample_dim) if (hasattr(self.model, 'engine') and predictors.ndim == 4 and
                                                not self.is_recurrent) else None
        if eng is not None and eng.can_rollout():
            series = self.model.rollout_host(predictors, steps)
        else:
            series = np.full((out_steps,) + predictors.shape, np.nan, dtype=np.float32)
            p = predictors.copy()
            for t in range(steps):
                if kwargs.get('verbose', 0) > 0:
                    print('Prediction step %d/%d' % (t + 1, steps))
                result = self.predict(p, **kwargs)
                p[:] = result[:] if self._n_steps == 1 else result[-1]
                series[t * self._n_steps:(t + 1) * self._n_steps, ...] = np.stack(result, axis=0)
        return _series_reshape(series, out_steps, sample_dim, self.time_dim, feature_shape, keep_time_dim)

    def evaluate(self, predictors, targets, **kwargs):
        return self.model.evaluate(predictors, targets, **kwargs)
