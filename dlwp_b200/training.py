"""
Training entry points of the front-end Model: `fit_generator`, `fit`, `train_on_batch`, `evaluate` with the Keras call
signatures the reference uses (DLWP/model/models.py:188-228, :384-402, :303-317; examples/train_functional.py:300-316).

Per batch (what keras does inside fit_generator): forward through the (possibly unrolled, shared-weight) net, loss =
sum_k loss_weights[k] * mse_k, backward, Adam -- all in libdlwp_b200 (dlwp_train_step / dlwp_train_adam).  With
torch.distributed initialised (one process per GPU) the flat gradient buffer is all-reduced between the two: plain data
parallelism, the multi-GPU mode of the reference's `multi_gpu_model` (models.py:104-109) without its CPU-hosted weights.
"""

import numpy as np

from .keras.engine import History
from .keras.losses import mean_squared_error


def _loss_is_mse(loss):
    if loss in ('mse', 'MSE', 'mean_squared_error') or loss is mean_squared_error:
        return True
    # DLWP.custom.latitude_weighted_loss(mean_squared_error, lats, output_shape, ...): an MSE on latitude-scaled tensors
    return getattr(loss, 'base_loss', None) is mean_squared_error and hasattr(loss, 'weights')


_ACC_REGULARIZE = {None: 0, 'mse': 1, 'mae': 2}


def _loss_is_acc(loss):
    """DLWP.custom.anomaly_correlation_loss(...) / DLWP.custom.acc_loss (custom.py:1036-1093)."""
    if loss == 'acc_loss':
        return True
    return all(hasattr(loss, a) for a in ('mean', 'regularize_mean', 'reverse')) and getattr(loss, '__name__', '') == 'acc_loss'


def _set_acc_loss(eng, loss):
    from . import _native as nat
    if loss == 'acc_loss':
        from .custom import acc_loss as loss
    if loss.regularize_mean not in _ACC_REGULARIZE:
        raise NotImplementedError("anomaly_correlation_loss(regularize_mean=%r): the device loss has None, 'mse' and 'mae'"
                                  % (loss.regularize_mean,))
    mean = None if loss.mean is None else np.ascontiguousarray(loss.mean, np.float32)
    nat.check(nat.lib().dlwp_train_loss_kind(eng.plan, 1, _ACC_REGULARIZE[loss.regularize_mean], 1 if loss.reverse else 0,
                                             None if mean is None else mean.ctypes.data, 0 if mean is None else mean.size),
              'dlwp_train_loss_kind')


def _loss_weight_map(model, eng):
    """(H, W) weight map of a latitude-weighted loss, or None."""
    loss = model.loss[0] if isinstance(model.loss, (list, tuple)) else model.loss
    w = getattr(loss, 'weights', None)
    if w is None or not hasattr(loss, 'base_loss'):
        return None
    H, W = eng.out_phys[0][1], eng.out_phys[0][2]
    w = np.asarray(w, np.float32)
    if w.size == 1:
        return None
    return np.ascontiguousarray(np.broadcast_to(w.reshape(w.shape[-2:]) if w.ndim >= 2 else w.reshape(H, 1), (H, W)))


def _engine(model, batch):
    from .engine import CompiledNet
    eng = getattr(model, '_train_engine', None)
    if eng is None or eng.max_batch < batch:
        state = None
        if eng is not None:
            state = eng.adam_state()          # a larger batch arrived: the Adam moments and step count move to the new plan
            eng.close()
        eng = CompiledNet(model, batch, force_ffma=True)
        if eng.max_batch < batch:
            raise MemoryError('training batch %d does not fit the device' % batch)
        if state is not None:
            eng.set_adam_state(*state)
        wmap = _loss_weight_map(model, eng)
        if wmap is not None:
            from . import _native as nat
            nat.check(nat.lib().dlwp_train_loss_weights(eng.plan, wmap.ctypes.data, wmap.size), 'dlwp_train_loss_weights')
        first = model.loss[0] if isinstance(model.loss, (list, tuple)) else model.loss
        if _loss_is_acc(first):
            _set_acc_loss(eng, first)
        model._train_engine = eng
    return eng


def _as_list(y, n):
    ys = list(y) if isinstance(y, (list, tuple)) else [y]
    if len(ys) != n:
        raise ValueError('Error when checking model target: expected %d target arrays but got %d' % (n, len(ys)))
    return ys


def _check_compiled(model):
    if model._compile_kwargs is None:
        raise RuntimeError('You must compile your model before using it.')
    loss = model.loss
    losses = loss if isinstance(loss, (list, tuple)) else [loss]
    if not (all(_loss_is_mse(l) for l in losses) or (all(_loss_is_acc(l) for l in losses) and
                                                     all(l is losses[0] or l == losses[0] for l in losses))):
        raise NotImplementedError('dlwp_b200 trains with loss="mse", DLWP.custom.latitude_weighted_loss(mean_squared_error) '
                                  'or one DLWP.custom.anomaly_correlation_loss for all outputs')


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except ImportError:
        pass
    return None


_native_comm = {}


def _allreduce_gradients(eng, dist):
    """Mean of the flat gradient buffer over the data-parallel ranks.  NCCL process groups: one ncclAllReduce inside the
    library on its own communicator (created once per process, dlwp_train_allreduce), on the stream of the backward;
    other back ends: torch.distributed on a tensor view of the buffer."""
    if dist.get_backend() == 'nccl':
        from . import _native as nat
        from .parallel import create_native_comm
        key = (dist.get_rank(), dist.get_world_size())
        if key not in _native_comm:
            _native_comm.clear()
            _native_comm[key] = create_native_comm(dist, *key)
        nat.check(nat.lib().dlwp_train_allreduce(eng.plan, _native_comm[key], eng._stream()), 'dlwp_train_allreduce')
        return
    g = eng.grad_tensor()
    dist.all_reduce(g)
    g.div_(dist.get_world_size())


def _step(model, x, y, train):
    import torch
    _check_compiled(model)
    def dev(a):          # CUDA tensors (DeviceSeriesGenerator) pass through; numpy batches are copied to the device
        if isinstance(a, torch.Tensor):
            return a.to(device='cuda', dtype=torch.float32).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    xd = dev(x)
    eng = _engine(model, xd.shape[0])
    ys = _as_list(y, eng.n_outputs)
    yd = [dev(v).reshape((xd.shape[0],) + p) for v, p in zip(ys, eng.out_phys)]
    lw = model.loss_weights
    losses, maes = eng.train_step(xd, yd, lw, backward=train)
    if train:
        opt = model.optimizer
        dist = _dist()
        if dist is not None:
            _allreduce_gradients(eng, dist)
        if opt.__class__.__name__ != 'Adam':
            raise NotImplementedError('dlwp_b200 implements the Adam update (the optimizer of every DLWP example)')
        if getattr(opt, 'amsgrad', False):
            raise NotImplementedError('Adam(amsgrad=True) is not implemented')
        penalty = eng.regularize()            # kernel_regularizer / bias_regularizer: gradient term + loss penalty
        lr = opt.lr * (1. / (1. + opt.decay * opt.iterations))
        eng.adam(lr, opt.beta_1, opt.beta_2, opt.epsilon)
        opt.iterations += 1
    else:
        penalty = eng.regularization_penalty()
    w = [1.0] * len(losses) if lw is None else list(lw)
    logs = {'loss': float(sum(wi * li for wi, li in zip(w, losses))) + penalty}
    if len(losses) > 1:
        for k, li in enumerate(losses):
            logs['output_%d_loss' % k] = float(li)
    if any(m in ('mae', 'mean_absolute_error') for m in (model.metrics or [])):
        if len(maes) == 1:
            logs['mean_absolute_error'] = float(maes[0])
        else:
            for k, m in enumerate(maes):
                logs['output_%d_mean_absolute_error' % k] = float(m)
    return logs


def train_on_batch(model, x, y, **kwargs):
    logs = _step(model, x, y, True)
    model._train_engine.pull_weights()
    return logs['loss'] if len(logs) == 1 else [logs[k] for k in logs]


def _iter_batches(data, steps=None, batch_size=32):
    """(x, y) tuple of arrays or a keras Sequence -> batches."""
    if isinstance(data, (tuple, list)):
        x, y = data[0], data[1]
        n = len(x)
        for s in range(0, n, batch_size):
            yb = [v[s:s + batch_size] for v in y] if isinstance(y, (list, tuple)) else y[s:s + batch_size]
            yield x[s:s + batch_size], yb
        return
    n = len(data) if steps is None else min(steps, len(data))
    for i in range(n):
        b = data[i]
        yield b[0], b[1]


def evaluate(model, x=None, y=None, batch_size=32, verbose=0, steps=None, **kwargs):
    data = x if y is None else (x, y)
    totals, count = {}, 0
    for xb, yb in _iter_batches(data, steps, batch_size):
        logs = _step(model, xb, yb, False)
        n = len(xb)
        for k, v in logs.items():
            totals[k] = totals.get(k, 0.0) + v * n
        count += n
    out = [totals[k] / max(count, 1) for k in totals]
    return out[0] if len(out) == 1 else out


def fit_generator(model, generator, steps_per_epoch=None, epochs=1, verbose=1, callbacks=None, validation_data=None,
                  validation_steps=None, class_weight=None, max_queue_size=10, workers=1, use_multiprocessing=False,
                  shuffle=True, initial_epoch=0, **kwargs):
    _check_compiled(model)
    history = History()
    cbs = [history] + list(callbacks or [])
    for cb in cbs:
        if hasattr(cb, 'set_model'):
            cb.set_model(model)
        if hasattr(cb, 'set_params'):
            cb.set_params({'epochs': epochs, 'steps': steps_per_epoch, 'verbose': verbose})
    model.stop_training = False
    model.history = history
    for cb in cbs:
        cb.on_train_begin()
    for epoch in range(initial_epoch, epochs):
        for cb in cbs:
            cb.on_epoch_begin(epoch)
        totals, count = {}, 0
        for bi, (xb, yb) in enumerate(_iter_batches(generator, steps_per_epoch)):
            for cb in cbs:
                cb.on_batch_begin(bi)
            logs = _step(model, xb, yb, True)
            n = len(xb)
            for k, v in logs.items():
                totals[k] = totals.get(k, 0.0) + v * n
            count += n
            for cb in cbs:
                cb.on_batch_end(bi, dict(logs, batch=bi, size=n))
        epoch_logs = {k: v / max(count, 1) for k, v in totals.items()}
        model._train_engine.pull_weights()
        if validation_data is not None:
            val = evaluate(model, validation_data, steps=validation_steps)
            vals = val if isinstance(val, list) else [val]
            for k, v in zip(list(epoch_logs.keys()), vals):
                epoch_logs['val_' + k] = v
        if verbose:
            print('Epoch %d/%d - ' % (epoch + 1, epochs) + ' - '.join('%s: %.6f' % kv for kv in epoch_logs.items()))
        for cb in cbs:
            cb.on_epoch_end(epoch, epoch_logs)
        if hasattr(generator, 'on_epoch_end'):
            generator.on_epoch_end()
        if model.stop_training:
            break
    for cb in cbs:
        cb.on_train_end()
    model._train_engine.pull_weights()
    return history


def fit(model, x=None, y=None, batch_size=32, epochs=1, verbose=1, callbacks=None, validation_data=None, shuffle=True,
        **kwargs):
    from .model.generators import ArrayDataGenerator
    gen = ArrayDataGenerator(x, y, batch_size=batch_size or 32, shuffle=shuffle)
    val = None
    if validation_data is not None:
        val = ArrayDataGenerator(validation_data[0], validation_data[1], batch_size=batch_size or 32)
    return fit_generator(model, gen, epochs=epochs, verbose=verbose, callbacks=callbacks, validation_data=val)
