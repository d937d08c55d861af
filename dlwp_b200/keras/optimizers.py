"""keras.optimizers subset: hyper-parameter holders; the update rule runs in dlwp_b200.training."""


class Optimizer(object):
    def __init__(self, **kwargs):
        self.iterations = 0
        self.decay = float(kwargs.pop('decay', 0.0))


class Adam(Optimizer):
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0., amsgrad=False, **kwargs):
        super(Adam, self).__init__(decay=decay, **kwargs)
        self.lr, self.beta_1, self.beta_2 = float(lr), float(beta_1), float(beta_2)
        self.epsilon = 1e-7 if epsilon is None else float(epsilon)  # K.epsilon()
        self.amsgrad = amsgrad


class SGD(Optimizer):
    def __init__(self, lr=0.01, momentum=0., decay=0., nesterov=False, **kwargs):
        super(SGD, self).__init__(decay=decay, **kwargs)
        self.lr, self.momentum, self.nesterov = float(lr), float(momentum), bool(nesterov)


def get(identifier):
    if isinstance(identifier, Optimizer):
        return identifier
    if isinstance(identifier, str):
        table = {'adam': Adam, 'sgd': SGD}
        if identifier.lower() in table:
            return table[identifier.lower()]()
    raise ValueError('Could not interpret optimizer identifier: ' + str(identifier))
