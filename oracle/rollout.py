"""
Restatement of the two rollout drivers.  TEST INFRASTRUCTURE (see oracle/__init__.py).

``predict_fn(p)`` plays the role of ``self.predict(p, **kwargs)`` (DLWP/model/models.py:230-245, 404-412): it maps an
``(N, ...)`` float array to the model output (a list of arrays for a multi-output functional model).
"""

import math

import numpy as np


def neuralnet_predict_timeseries(predict_fn, predictors, time_steps, time_dim=1, is_recurrent=False,
                                 step_sequence=False, keep_time_dim=False, dtype=np.float32):
    """
    DLWPNeuralNet.predict_timeseries, DLWP/model/models.py:247-301.

    * 265-269: ``time_steps`` -> int; < 1 raises ValueError; unless step_sequence the model runs
      ceil(time_steps / time_dim) times (the result is NOT truncated to time_steps).
    * 270-276: float32 NaN-filled buffer (steps,)+predictors.shape; private copy of the input.
    * 277-293: the feedback loop; the step_sequence branch (280-290) drops the oldest time slice and appends the first
      predicted slice.
    * 294-300: (steps, N, time_dim, C', ...) reshape; without keep_time_dim the time_dim axis is folded into the leading
      forecast-hour axis (or slice 0 is taken for step_sequence).

    ``dtype`` is float32 in the reference; tests pass float64 to get a drift-free tier-0 series.
    """
    time_steps = int(time_steps)
    if time_steps < 1:
        raise ValueError("time_steps must be an int > 0")
    if not step_sequence:
        time_steps = int(math.ceil(1. * time_steps / time_dim))
    series = np.full((time_steps,) + predictors.shape, np.nan, dtype=dtype)
    p = predictors.copy()
    n = p.shape[0]
    feature_shape = p.shape[2:] if is_recurrent else p.shape[1:]
    for t in range(time_steps):
        if step_sequence:
            pr = predict_fn(p)
            pr_shape = pr.shape[:]
            if not is_recurrent:
                pr = pr.reshape((n, time_dim, -1) + feature_shape[1:])
                p = p.reshape((n, time_dim, -1) + feature_shape[1:])
            p = np.concatenate((p[:, 1:], pr[:, [0]]), axis=1)
            if not is_recurrent:
                p = p.reshape(predictors.shape)
                pr = pr.reshape(pr_shape)
            series[t, ...] = pr
        else:
            p = 1. * predict_fn(p)
            series[t, ...] = p
    series = series.reshape((time_steps, n, time_dim, -1) + feature_shape[1:])
    if not keep_time_dim:
        if step_sequence:
            series = series[:, :, 0]
        else:
            series = series.transpose((0, 2, 1) + tuple(range(3, 3 + len(feature_shape))))
            series = series.reshape((time_steps * time_dim, n, -1) + feature_shape[1:])
    return series


def functional_predict_timeseries(predict_fn, predictors, time_steps, n_steps=1, time_dim=1, is_recurrent=False,
                                  keep_time_dim=False, dtype=np.float32):
    """
    DLWPFunctional.predict_timeseries, DLWP/model/models.py:414-452.

    * 427-431: steps = ceil(time_steps / n_steps / time_dim); out_steps = steps * n_steps (n_steps = len(model.outputs),
      models.py:364).
    * 439-447: each iteration feeds back the LAST unrolled output (443-446) and stores all of them (447).  For
      n_steps == 1 the result is a bare (N, ...) array and np.stack(result, axis=0) rebuilds it unchanged (SURVEY.md
      Appendix A.7(i)).
    * 448-451: same reshape/transposition rule as the Sequential driver.
    """
    time_steps = int(time_steps)
    if time_steps < 1:
        raise ValueError("time_steps must be an int > 0")
    steps = int(math.ceil(time_steps / n_steps / time_dim))
    out_steps = steps * n_steps
    series = np.full((out_steps,) + predictors.shape, np.nan, dtype=dtype)
    p = predictors.copy()
    n = p.shape[0]
    feature_shape = p.shape[2:] if is_recurrent else p.shape[1:]
    for t in range(steps):
        result = predict_fn(p)
        if n_steps == 1:
            p[:] = result[:]
        else:
            p[:] = result[-1]
        series[t * n_steps:(t + 1) * n_steps, ...] = np.stack(result, axis=0)
    series = series.reshape((out_steps, n, time_dim, -1) + feature_shape[1:])
    if not keep_time_dim:
        series = series.transpose((0, 2, 1) + tuple(range(3, 3 + len(feature_shape))))
        series = series.reshape((out_steps * time_dim, n, -1) + feature_shape[1:])
    return series
