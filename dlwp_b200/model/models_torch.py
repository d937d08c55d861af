"""
`DLWP.model.DLWPTorchNN` (reference DLWP/model/models_torch.py:28-407): the torch.nn twin of `DLWPNeuralNet` -- a network
given as `(torch_layer_name, args, kwargs)` tuples, `predict` / `predict_timeseries` / `fit_generator` / `evaluate` with the
reference's signatures -- with the arithmetic on the B200 path.

The reference keeps `self.layers` (torch modules) and runs them one by one on its `device`.  Here the torch modules are
kept too, on the CPU, as the user-visible PARAMETER HOLDERS (`dlwp.layers[i].weight`, `state_dict`, `reset`), while every
forward runs through the same engine as `DLWPNeuralNet`: the layer list is translated once the input shape is known
(torch layers do not declare it) into the equivalent Keras-style stack -- CircularPad2d -> PeriodicPadding2D, ZeroPad2d ->
ZeroPadding2D, ReflectionPad2d -> TFPadding2D(REFLECT), ReplicationPad2d -> FillPadding2D, Conv2d (stride 1, groups 1;
`padding=` / `padding_mode=` become an explicit padding layer) -> Conv2D, MaxPool2d(2) -> MaxPooling2D, Upsample(2, nearest)
-> UpSampling2D -- and weights travel as (Cout, Cin, kh, kw) <-> Keras (kh, kw, Cin, Cout).  `predict_timeseries` is then
ONE device-resident rollout (models_torch.py:322-376 restated by `DLWPNeuralNet.predict_timeseries`, whose output rule is
the same).  Training maps `optimizer='Adam'` + `loss='MSELoss'` onto `dlwp_train_step` / `dlwp_train_adam`; anything else
raises NotImplementedError.
"""

import time

import numpy as np

from .. import util
from .models import DLWPNeuralNet

_ACTIVATIONS = {'tanh': 'tanh', 'relu': 'relu'}       # torch.nn.functional names the conv epilogues implement


def _pad4(padding):
    """torch's (left, right, top, bottom) or int -> Keras ((top, bottom), (left, right))."""
    if isinstance(padding, int):
        return ((padding, padding), (padding, padding))
    pl, pr, pt, pb = [int(v) for v in padding]
    return ((pt, pb), (pl, pr))


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else tuple(int(a) for a in v)


class DLWPTorchNN(object):
    def __init__(self, is_convolutional=False, is_recurrent=False, time_dim=1, scaler_type='StandardScaler',
                 scale_targets=True, apply_same_y_scaling=True, impute_missing=False):
        # the twin that owns scalers / imputers and runs the engine (constructor validation is shared: models_torch.py:50-52)
        self._net = DLWPNeuralNet(is_convolutional=is_convolutional, is_recurrent=is_recurrent, time_dim=time_dim,
                                  scaler_type=scaler_type, scale_targets=scale_targets,
                                  apply_same_y_scaling=apply_same_y_scaling, impute_missing=impute_missing)
        self.is_convolutional, self.is_recurrent, self.time_dim = is_convolutional, is_recurrent, time_dim
        self.scaler_type, self.scale_targets, self.apply_same_y_scaling = scaler_type, scale_targets, apply_same_y_scaling
        self.impute = impute_missing
        self.model = None
        self.optimizer = None
        self.loss = None
        self.metric = None
        self.layers = []
        self.activations = []
        self.history = {}
        self._specs = None
        self._built_for = None          # input shape the engine-side model was built for
        self._versions = None           # torch parameter versions at the last weight sync

    # scalers / imputers live on the twin (same code as the reference's, models_torch.py:163-232 == models.py:114-194)
    scaler = property(lambda self: self._net.scaler)
    scaler_y = property(lambda self: self._net.scaler_y)
    imputer = property(lambda self: self._net.imputer)
    imputer_y = property(lambda self: self._net.imputer_y)

    def scaler_fit(self, X, y, **kwargs):
        return self._net.scaler_fit(X, y, **kwargs)

    def scaler_transform(self, X, y=None):
        return self._net.scaler_transform(X, y)

    def imputer_fit(self, X, y):
        return self._net.imputer_fit(X, y)

    def imputer_transform(self, X, y=None):
        return self._net.imputer_transform(X, y)

    def init_fit(self, predictors, targets):
        return self._net.init_fit(predictors, targets)

    # -- construction (models_torch.py:77-153) -----------------------------------------------------------------------------
    def build_model(self, layers, optimizer, loss, optimizer_kwargs=None, loss_kwargs=None, metric='L1Loss',
                    metric_kwargs=None):
        from torch import nn
        if type(layers) not in [list, tuple]:
            raise TypeError("'layers' argument must be a tuple")
        layers = [l for l in layers]
        for l, layer in enumerate(layers):
            if type(layer) not in [list, tuple]:
                raise TypeError("each element of 'layers' must be a tuple")
            if len(layer) != 3:
                raise ValueError("each layer must be specified by three elements (name, args, kwargs)")
            if layer[1] is None:
                layer = [layer[0], (), layer[2]]
            if type(layer[1]) is not tuple:
                raise TypeError("the 'args' element of layer %d must be a tuple" % l)
            if layer[2] is None:
                layer = [layer[0], layer[1], {}]
            if type(layer[2]) is not dict:
                raise TypeError("the 'kwargs' element of layer %d must be a dict" % l)
            layers[l] = [layer[0], layer[1], dict(layer[2])]
        for name, value in (('optimizer_kwargs', optimizer_kwargs), ('loss_kwargs', loss_kwargs),
                            ('metric_kwargs', metric_kwargs)):
            if value is not None and value != {} and not isinstance(value, dict):
                raise TypeError("'%s' must be a dict" % name)
        optimizer_kwargs, loss_kwargs, metric_kwargs = optimizer_kwargs or {}, loss_kwargs or {}, metric_kwargs or {}

        self.model = nn.Module()
        self.layers, self.activations, self._specs = [], [], []
        for name, args, kwargs in layers:
            try:
                layer_class = util.get_from_class('torch.nn', name)
            except (ImportError, AttributeError):
                layer_class = util.get_from_class('DLWP.custom', name)
            act = kwargs.pop('activation', None)
            self.activations.append(util.get_from_class('torch.nn.functional', act) if act is not None else None)
            self.layers.append(layer_class(*args, **kwargs))
            self._specs.append((name, act))
        for l, layer in enumerate(self.layers):
            setattr(self.model, 'layer%d' % l, layer)
        self.model.forward = self._forward
        self.keras_layers(None)          # refuse untranslatable layers now, not at the first predict
        self.loss = util.get_from_class('torch.nn', loss)(**loss_kwargs)
        self.optimizer = util.get_from_class('torch.optim', optimizer)(self.model.parameters(), **optimizer_kwargs)
        self.metric = util.get_from_class('torch.nn', metric)(**metric_kwargs)
        self._built_for = None

    def _forward(self, x):
        """The reference runs the torch modules here (models_torch.py:155-160).  In this package they only HOLD the parameters:
        there is no CPU forward -- every forward goes through the GPU engine."""
        raise RuntimeError('DLWPTorchNN.model holds the parameters only; use predict / predict_timeseries (GPU engine)')

    # -- translation -----------------------------------------------------------------------------------------------------------
    def keras_layers(self, input_shape):
        """The `(name, args, kwargs)` tuples of the equivalent DLWPNeuralNet.build_model call (channels_first)."""
        from torch import nn
        cf = 'channels_first'
        out = []

        def add(name, args, kwargs):
            kwargs = dict(kwargs, data_format=cf)
            if not out and input_shape is not None:
                kwargs['input_shape'] = tuple(input_shape)
            out.append((name, args, kwargs))

        for (name, act), layer in zip(self._specs, self.layers):
            if act is not None and not isinstance(layer, nn.Conv2d):
                raise NotImplementedError("an 'activation' on a %s layer has no fused equivalent" % name)
            if isinstance(layer, nn.Conv2d):
                if act is not None and act not in _ACTIVATIONS:
                    raise NotImplementedError('activation %r' % (act,))
                if _pair(layer.stride) != (1, 1) or layer.groups != 1:
                    raise NotImplementedError('Conv2d with stride != 1 or groups != 1')
                pad = layer.padding
                if isinstance(pad, str):
                    if pad != 'valid':
                        raise NotImplementedError("Conv2d(padding=%r): give the padding as integers" % (pad,))
                    pad = (0, 0)
                ph, pw = _pair(pad)
                if ph or pw:
                    mode = {'zeros': 'ZeroPadding2D', 'circular': 'PeriodicPadding2D'}.get(layer.padding_mode)
                    if mode is None:
                        raise NotImplementedError('Conv2d(padding_mode=%r)' % (layer.padding_mode,))
                    add(mode, (((ph, ph), (pw, pw)),), {})
                add('Conv2D', (layer.out_channels, _pair(layer.kernel_size)),
                    {'dilation_rate': _pair(layer.dilation), 'activation': _ACTIVATIONS.get(act, 'linear'),
                     'use_bias': layer.bias is not None, 'padding': 'valid'})
            elif isinstance(layer, nn.ZeroPad2d):
                add('ZeroPadding2D', (_pad4(layer.padding),), {})
            elif hasattr(nn, 'CircularPad2d') and isinstance(layer, nn.CircularPad2d):
                add('PeriodicPadding2D', (_pad4(layer.padding),), {})
            elif isinstance(layer, nn.ReflectionPad2d):
                add('TFPadding2D', (_pad4(layer.padding),), {'mode': 'REFLECT'})
            elif isinstance(layer, nn.ReplicationPad2d):
                add('FillPadding2D', (_pad4(layer.padding),), {})
            elif isinstance(layer, nn.MaxPool2d):
                if _pair(layer.kernel_size) != (2, 2) or _pair(layer.stride or layer.kernel_size) != (2, 2) or \
                        _pair(layer.padding) != (0, 0) or _pair(layer.dilation) != (1, 1) or layer.ceil_mode:
                    raise NotImplementedError('MaxPool2d other than kernel 2, stride 2')
                add('MaxPooling2D', (2,), {})
            elif isinstance(layer, nn.Upsample):
                sf = layer.scale_factor
                if layer.mode != 'nearest' or sf is None or tuple(np.atleast_1d(sf).astype(int).tolist()) not in ((2,), (2, 2)):
                    raise NotImplementedError("Upsample other than scale_factor=2, mode='nearest'")
                add('UpSampling2D', (2,), {})
            else:
                raise NotImplementedError('torch layer %s has no equivalent on the B200 rollout path' % name)
        return tuple(out)

    def _conv_modules(self):
        from torch import nn
        return [m for m in self.layers if isinstance(m, nn.Conv2d)]

    def keras_weights(self):
        """Weights of the torch Conv2d modules in the order and layout of `keras Model.get_weights()`."""
        out = []
        for m in self._conv_modules():
            out.append(np.ascontiguousarray(np.transpose(m.weight.detach().cpu().numpy(), (2, 3, 1, 0)), np.float32))
            if m.bias is not None:
                out.append(m.bias.detach().cpu().numpy().astype(np.float32))
        return out

    def _param_versions(self):
        return [p._version for m in self._conv_modules() for p in m.parameters()]

    def _ensure(self, input_shape):
        """Build (once per input shape) the engine-side model and push the torch parameters when they have changed."""
        if self.model is None:
            raise RuntimeError('build_model has not been called')
        input_shape = tuple(int(v) for v in input_shape)
        if self._built_for != input_shape:
            self._net.build_model(self.keras_layers(input_shape), loss='mse', optimizer=self._keras_optimizer(),
                                  metrics=['mae'])
            self._built_for, self._versions = input_shape, None
        versions = self._param_versions()
        if versions != self._versions:
            self._net.model.set_weights(self.keras_weights())
            self._versions = versions

    def _pull(self):
        """Engine weights (after training steps) back into the torch parameter holders."""
        import torch
        ws = iter(self._net.model.get_weights())
        with torch.no_grad():
            for m in self._conv_modules():
                m.weight.copy_(torch.from_numpy(np.ascontiguousarray(np.transpose(next(ws), (3, 2, 0, 1)))))
                if m.bias is not None:
                    m.bias.copy_(torch.from_numpy(np.ascontiguousarray(next(ws))))
        self._versions = self._param_versions()

    def _keras_optimizer(self):
        from ..keras import optimizers
        opt = self.optimizer
        if opt is None or opt.__class__.__name__ != 'Adam':
            return 'adam'                # prediction only needs a compiled model; training checks in fit_generator
        g = opt.param_groups[0]
        return optimizers.Adam(lr=g['lr'], beta_1=g['betas'][0], beta_2=g['betas'][1], epsilon=g['eps'])

    # -- prediction (models_torch.py:301-376) ----------------------------------------------------------------------------------
    def predict(self, predictors):
        predictors = np.asarray(predictors)
        self._ensure(predictors.shape[1:])
        return self._net.predict(predictors)

    def predict_timeseries(self, predictors, time_steps, step_sequence=False, keep_time_dim=False, verbose=0):
        predictors = np.asarray(predictors)
        if int(time_steps) < 1:
            raise ValueError("time_steps must be an int > 0")
        self._ensure(predictors.shape[1:])
        return self._net.predict_timeseries(predictors, time_steps, step_sequence=step_sequence,
                                            keep_time_dim=keep_time_dim, verbose=verbose)

    # -- training (models_torch.py:234-299, 378-397) ---------------------------------------------------------------------------
    def _check_trainable(self):
        if self.optimizer.__class__.__name__ != 'Adam' or self.loss.__class__.__name__ != 'MSELoss' or \
                getattr(self.loss, 'reduction', 'mean') != 'mean':
            raise NotImplementedError("dlwp_b200 trains DLWPTorchNN with optimizer='Adam' and loss='MSELoss'")
        g = self.optimizer.param_groups[0]
        if g.get('weight_decay', 0) or g.get('amsgrad', False):
            raise NotImplementedError('Adam(weight_decay / amsgrad) is not implemented')
        if self.metric.__class__.__name__ != 'L1Loss':
            raise NotImplementedError("the error metric of the device training step is 'L1Loss'")

    def _torch_adam_step_constants(self):
        """torch.optim.Adam adds eps to sqrt(v / (1 - b2^t)), Keras 2.2 Adam (what dlwp_train_adam implements) to sqrt(v) with
        the bias corrections folded into the step size: the two coincide when Keras' epsilon is eps * sqrt(1 - b2^t).  The
        learning rate is re-read every step (a torch lr scheduler edits param_groups)."""
        g, opt = self.optimizer.param_groups[0], self._net.model.optimizer
        opt.lr, opt.beta_1, opt.beta_2 = float(g['lr']), float(g['betas'][0]), float(g['betas'][1])
        opt.epsilon = float(g['eps']) * float(np.sqrt(1.0 - opt.beta_2 ** (opt.iterations + 1)))

    def _batch(self, generator, b):
        p, t = generator[b]
        return np.asarray(p, np.float32), np.asarray(t, np.float32)

    def fit_generator(self, generator, epochs=1, min_epochs=None, validation_generator=None, early_stop=None,
                      lr_schedule=None, verbose=0):
        from .. import training
        self._check_trainable()
        self.history['loss'], self.history['error'] = [], []
        if validation_generator is not None:
            self.history['val_loss'], self.history['val_error'] = [], []
        elif lr_schedule is not None:
            print("Warning: learning rate scheduler 'lr_sched' needs validation data; disabling")
        n_d = len(generator)
        for epoch in range(epochs):
            if verbose > 0:
                print('\nEpoch %d/%d' % (epoch + 1, epochs))
            epoch_start = time.time()
            running_loss = running_error = 0.0
            for b in range(n_d):
                p, t = self._batch(generator, b)
                self._ensure(p.shape[1:])
                self._torch_adam_step_constants()
                logs = training._step(self._net.model, p, t, True)
                running_loss = (b * running_loss + logs['loss']) / (b + 1)
                running_error = (b * running_error + logs['mean_absolute_error']) / (b + 1)
                if verbose > 1:
                    print('%d/%d loss: %0.4f - error: %0.4f' % (b + 1, n_d, running_loss, running_error), end='\r')
            print_line = ''
            self.history['loss'].append(running_loss)
            self.history['error'].append(running_error)
            if verbose > 0:
                print_line += ' - loss: %0.4f - error: %0.4f' % (running_loss, running_error)
            if validation_generator is not None:
                running_loss = running_error = 0.0
                for b in range(len(validation_generator)):
                    p, t = self._batch(validation_generator, b)
                    logs = training._step(self._net.model, p, t, False)
                    running_loss = (b * running_loss + logs['loss']) / (b + 1)
                    running_error = (b * running_error + logs['mean_absolute_error']) / (b + 1)
                self.history['val_loss'].append(running_loss)
                self.history['val_error'].append(running_error)
                if verbose > 0:
                    print_line += ' - val_loss: %0.4f - val_error: %0.4f' % (running_loss, running_error)
                if early_stop is not None:
                    if min_epochs is not None and epoch > min_epochs + early_stop:
                        if epoch - np.argmin(self.history['val_loss']) == early_stop:
                            if verbose > 0:
                                print('\nval_loss stopped improving; ending fit')
                            break
                if lr_schedule is not None:
                    lr_schedule.step(running_loss)
            if verbose > 0:
                print('%d/%d - time: %0.2f s' % (n_d, n_d, time.time() - epoch_start) + print_line, end='')
        if verbose > 0:
            print('')
        if getattr(self._net.model, '_train_engine', None) is not None:
            self._net.model._train_engine.pull_weights()
        self._pull()
        return self.history

    def evaluate(self, predictors, targets):
        from .. import training
        self._check_trainable()
        if self.impute:
            predictors, targets = self.imputer_transform(predictors, targets)
        p, t = self.scaler_transform(predictors, targets)
        p, t = np.asarray(p, np.float32), np.asarray(t, np.float32)
        self._ensure(p.shape[1:])
        logs = training._step(self._net.model, p, t, False)
        return logs['loss'], logs['mean_absolute_error']

    def reset(self):
        """models_torch.py:399-407."""
        for c in self.model.named_children():
            try:
                c[1].reset_parameters()
            except AttributeError:
                print("warning: layer '%s' cannot be reset" % c[0])
        self._versions = None
