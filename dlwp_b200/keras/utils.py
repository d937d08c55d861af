"""keras.utils names the reference imports (DLWP/model/models.py:15, DLWP/model/generators.py)."""


class Sequence(object):
    def __getitem__(self, index):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    def on_epoch_end(self):
        pass


def multi_gpu_model(model, gpus=None, **kwargs):
    """keras.utils.multi_gpu_model as the reference uses it (DLWP/model/models.py:104-109, DLWP/util.py:176-183): a
    single-process replica per device, every batch split evenly over them, results concatenated on the host.  The replicas
    share the model's layers (one set of weights, pushed to each device's plan); `predict` / `predict_timeseries` run one
    host thread per device.  Training through a multi_gpu_model is not provided -- data-parallel training is one process
    per GPU (dlwp_b200/training.py: gradient all-reduce over NCCL)."""
    if gpus is None or isinstance(gpus, (list, tuple)):
        gpus = len(gpus) if gpus is not None else 1
    if int(gpus) <= 1:
        raise ValueError('For multi-gpu usage to be effective, call `multi_gpu_model` with `gpus >= 2`. Received: '
                         '`gpus=%s`' % (gpus,))
    model._gpus = int(gpus)
    return model
