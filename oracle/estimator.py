"""
Restatement of `TimeSeriesEstimator.predict` (DLWP/model/extensions.py:136-303) and `insolation`
(DLWP/util.py:300-352) on plain numpy arrays.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference drives the loop through xarray (`reindex`, `.loc`), which is not available here or on the GPU box.  Pinning:
`insolation` against the reference function run from its source text (tests/golden/insolation.npz); the loop against the
reference's OWN `TimeSeriesEstimator.predict`, executed by tests/golden/make_golden.py:gen_estimator on a numpy stand-in for
the few xarray calls it makes (tests/golden/fake_xarray.py: `reindex` = exact label lookup with NaN fill, `.loc` = label ->
position, views for scalar labels) -- tests/golden/estimator.npz, six cases (tests/test_oracle_golden.py).  The stand-in is
ours, so the pin is on the reference's loop logic, not on xarray itself.  This file follows the cited lines with index
arithmetic in place of label lookups --

* `p_da.reindex(sample=r_da.sample)` (extensions.py:226): r_da.sample = p_da.sample + (es + interval - 1)·dt on an evenly
  spaced sample axis, i.e. new p[i] = old p[i + shift], NaN where i + shift runs past the data;
* `p_da.loc[{'varlev': ...}]` (extensions.py:242-252): positions of the named varlevs in the input / output selections.
"""

import numpy as np


def day_of_year(dates):
    """DLWP/util.py:300-302 for an array of numpy datetime64: fractional days since January 1st of the date's year."""
    d = np.asarray(dates, 'datetime64[s]')
    start = d.astype('datetime64[Y]').astype('datetime64[s]')
    return (d - start).astype(np.float64) / 3600. / 24.


def insolation(dates, lat, lon, S=1.):
    """DLWP/util.py:305-352: approximate top-of-atmosphere insolation (date, lat, lon), float32, negative values clipped."""
    lat, lon = np.asarray(lat, np.float64), np.asarray(lon, np.float64)
    if lat.ndim != lon.ndim:
        raise ValueError("'lat' and 'lon' must either both be 1d or both be 2d'")
    if lat.ndim == 2 and lat.shape != lon.shape:
        raise ValueError("shape mismatch between lat (%s) and lon (%s)" % (lat.shape, lon.shape))
    if lat.ndim == 1:
        lon, lat = np.meshgrid(lon, lat)
    eps = 23.4441 * np.pi / 180.
    ecc = 0.016715
    om = 282.7 * np.pi / 180.
    beta = np.sqrt(1 - ecc ** 2.)
    days = day_of_year(dates)
    lambda_m0 = ecc * (1. + beta) * np.sin(om)
    lambda_m = lambda_m0 + 2. * np.pi * (days - 80.5) / 365.
    lambda_ = lambda_m + 2. * ecc * np.sin(lambda_m - om)
    dec = np.arcsin(np.sin(eps) * np.sin(lambda_))
    h = 2 * np.pi * (days[:, None, None] + lon / 360.)
    rho = (1. - ecc ** 2.) / (1. + ecc * np.cos(lambda_ - om))
    lat = lat * np.pi / 180.
    sol = S * (np.sin(lat[None, ...]) * np.sin(dec[:, None, None]) -
               np.cos(lat[None, ...]) * np.cos(dec[:, None, None]) * np.cos(h)) * rho[:, None, None] ** -2.
    sol[sol < 0.] = 0.
    return sol.astype(np.float32)


def estimator_predict(predict_fn, p, steps, t_in, t_out, in_varlev, out_varlev, sample_times, dt, lat, lon,
                      interval=1, add_insolation=False, impute=False, prefer_first_times=True):
    """
    extensions.py:158-253 for a Sequential model.  p: (S, t_in * V_in, H, W) unscaled predictors as the generator returns
    them (`in_varlev` includes 'SOL' last when add_insolation).  predict_fn maps (S, t_in*V_in, H, W) to
    (S, t_out*V_out, H, W).  Returns (result (effective_steps, S, t_out, V_out, H, W), es, keep_inputs).
    """
    steps = int(steps)
    if steps < 1:
        raise ValueError('must use positive integer for steps')
    in_varlev, out_varlev = list(in_varlev), list(out_varlev)
    if t_out <= t_in:                                   # 159-163
        keep_inputs, es = True, t_out
    else:                                               # 164-171
        keep_inputs = False
        es = t_in if prefer_first_times else t_out
    effective_steps = int(np.ceil(steps / es))          # 172
    S = p.shape[0]
    H, W = p.shape[-2:]
    p_shape = p.shape
    p = np.array(p, np.float32).reshape(S, t_in, -1, H, W)      # 179
    times = np.asarray(sample_times).astype('datetime64[s]')
    dt = np.timedelta64(dt, 's') if not isinstance(dt, np.timedelta64) else dt.astype('timedelta64[s]')
    shared = [v for v in out_varlev if v in in_varlev]   # _outputs_in_inputs (extensions.py:95-97)
    idx_in = [in_varlev.index(v) for v in shared]
    idx_out = [out_varlev.index(v) for v in shared]
    p_mean = p.mean(axis=0) if impute else None          # 191-192
    shift = es + interval - 1
    V_out = len(out_varlev)
    result = np.full((effective_steps, S, t_out, V_out, H, W), np.nan, np.float32)
    for s in range(effective_steps):                     # 210-253
        r = np.asarray(predict_fn(p.reshape(p_shape)), np.float32).reshape(S, t_out, V_out, H, W)
        result[s] = r
        times = times + shift * dt                       # r_da.sample (218)
        q = np.full_like(p, np.nan)                      # reindex (226)
        if shift < S:
            q[:S - shift] = p[shift:]
        p = q
        if impute:                                       # 229-231
            p[-es:] = p_mean[None]
        if add_insolation:                               # 234-238
            k = in_varlev.index('SOL')
            for n in range(t_in):
                p[-es:, n, k] = insolation(times[-es:] + n * dt, lat, lon)
        if keep_inputs:                                  # 242-244
            p[:, t_in - es:, idx_in] = r[:, :, idx_out]
        elif prefer_first_times:                         # 246-248
            p[:, :, idx_in] = r[:, :t_in][:, :, idx_out]
        else:                                            # 250-252
            p[:, :, idx_in] = r[:, t_out - t_in:][:, :, idx_out]
    return result, es, keep_inputs


def estimator_series(result, steps, es, keep_inputs, keep_time_dim=False, prefer_first_times=True):
    """extensions.py:257-292 without the coordinates: keep_time_dim -> (f_hour, time, time_step, varlev, lat, lon), else the
    continuous series (f_hour, time, varlev, lat, lon) cut to `steps` forecast hours."""
    if keep_time_dim:
        return result
    if not keep_inputs and prefer_first_times:
        result = result[:, :, :es]
    result = result.transpose((0, 2, 1, 3, 4, 5))
    result = result.reshape((-1,) + result.shape[2:])
    return result[:steps]
