"""keras.regularizers.l2 (examples/train_functional.py:26).  Recorded on the layer; only the training path reads it."""


class L1L2(object):
    def __init__(self, l1=0., l2=0.):
        self.l1 = float(l1)
        self.l2 = float(l2)

    def get_config(self):
        return {'l1': self.l1, 'l2': self.l2}


def l2(l=0.01):
    return L1L2(l2=l)


def l1(l=0.01):
    return L1L2(l1=l)
