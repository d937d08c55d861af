"""keras.utils names the reference imports (DLWP/model/models.py:15, DLWP/model/generators.py)."""


class Sequence(object):
    def __getitem__(self, index):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    def on_epoch_end(self):
        pass


def multi_gpu_model(model, gpus=None, **kwargs):
    """The reference's only multi-device mechanism (single-process batch split, models.py:104-109).  dlwp_b200 scales
    with one process per GPU instead (dlwp_b200.parallel); inside one process this is the identity."""
    return model
