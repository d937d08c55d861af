"""
Tier-1 comparator and CPU timing stand-in: the oracle forward in torch-CPU fp32.  TEST INFRASTRUCTURE.

``KerasLikeModel.predict`` mimics what ``keras.Model.predict`` does around the graph on the hot path (SURVEY.md
Appendix A.5): split the batch into ``batch_size`` (default 32) chunks, run each, concatenate on the host.  Plugged into
oracle.rollout it is the "port" CPU baseline that bench.py reports: the reference's rollout loop around a fp32 CPU
forward with all host threads (``torch.set_num_threads(os.cpu_count())``).  It is FASTER than the reference's real
Keras/TF-1.x CPU path would be (no NCHW<->NHWC transposes, no session feed/fetch), so speed-ups quoted against it are
conservative.
"""

import os

import numpy as np
import torch


def use_all_cores():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    return n


class KerasLikeModel(object):
    def __init__(self, net):
        self.net = net
        self.n_outputs = getattr(net, 'n_outputs', 1)

    def predict(self, x, batch_size=32, verbose=0, steps=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        outs = None
        with torch.no_grad():
            for s in range(0, x.shape[0], batch_size):
                y = self.net.forward(torch.from_numpy(x[s:s + batch_size]))
                ys = y if isinstance(y, (list, tuple)) else [y]
                if outs is None:
                    outs = [[] for _ in ys]
                for o, v in zip(outs, ys):
                    o.append(v.numpy())
        outs = [np.concatenate(o, axis=0) for o in outs]
        return outs[0] if self.n_outputs == 1 else outs
