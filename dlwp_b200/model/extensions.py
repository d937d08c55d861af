"""
`DLWP.model.TimeSeriesEstimator` (reference DLWP/model/extensions.py:21-303): the caller of the rollout that handles
non-matching inputs and outputs -- outputs re-inserted among the inputs by variable/level name, inputs the model does not
predict taken from the known data of later samples, the insolation forcing (DLWP/util.py:305-352) recomputed for every new
time, fewer output than input time steps.

The reference walks an xarray DataArray through `reindex` / `.loc` on the HOST once per step (extensions.py:215-253).  Here
the predictor array lives on the GPU for the whole forecast: per step one model application (`CompiledNet.forward_device`),
one contiguous device copy for the sample shift, the insolation planes of the added times copied in, and one strided
channel copy (`dlwp_copy4d`) per run of re-inserted outputs -- nothing returns to the host until the series is complete.

xarray is not available offline (nor on the GPU box), so the estimator reads its metadata through plain attributes: any
generator object with the attributes of `dlwp_b200.model.ArraySeriesGenerator` works (a `SeriesDataGenerator` of the
reference has the same private names, extensions.py:37-135).  `predict` returns an `xarray.DataArray` when xarray can be
imported, else a `LabeledArray` (values + dims + coords) with the same dimension names and coordinate values.
"""

import ctypes

import numpy as np

from .. import _native as nat
from ..util import insolation
from .models import DLWPFunctional, DLWPNeuralNet
from .models_torch import DLWPTorchNN


class LabeledArray(object):
    """What `TimeSeriesEstimator.predict` returns where xarray is missing: `.values`, `.dims`, `.coords` (dict)."""

    def __init__(self, values, dims, coords):
        self.values, self.dims, self.coords = values, tuple(dims), dict(coords)
        self.shape = values.shape

    def __array__(self, dtype=None):
        return self.values if dtype is None else self.values.astype(dtype)


def _wrap(values, dims, coords):
    try:
        import xarray as xr
    except ImportError:
        return LabeledArray(values, dims, dict(zip(dims, coords)))
    return xr.DataArray(values, coords=coords, dims=dims)


class TimeSeriesEstimator(object):
    def __init__(self, model, generator):
        """extensions.py:28-135.  `model`: DLWPNeuralNet / DLWPFunctional; `generator`: see the module docstring."""
        if not isinstance(model, (DLWPNeuralNet, DLWPFunctional, DLWPTorchNN)):
            raise TypeError("'model' must be a valid instance of a DLWP model class")
        for attr in ('generate', 'convolution_shape', '_n_sample', 'sample_times', 'lat', 'lon'):
            if not hasattr(generator, attr):
                raise TypeError("'generator' must be a valid instance of a DLWP generator class (missing %r)" % attr)
        self.model = model
        self.generator = generator
        self._add_insolation = bool(getattr(generator, '_add_insolation', False))
        self._interval = int(getattr(generator, '_interval', 1))
        times = np.asarray(generator.sample_times).astype('datetime64[s]')
        self._dt = times[1] - times[0]
        self._input_sel = {'varlev': np.array(self._varlev_of(generator, generator._input_sel))}
        self._output_sel = {'varlev': np.array(self._varlev_of(generator, generator._output_sel))}
        self._outputs_in_inputs = {
            'varlev': np.array([v for v in self._output_sel['varlev'] if v in self._input_sel['varlev']])}
        if self._add_insolation:
            self._input_sel['varlev'] = np.concatenate([self._input_sel['varlev'], np.array(['SOL'])])
        self._input_time_steps = int(getattr(generator, '_input_time_steps', model.time_dim))
        self._output_time_steps = int(getattr(generator, '_output_time_steps', model.time_dim))

    # -- helpers ---------------------------------------------------------------------------------------------------------
    @staticmethod
    def _varlev_of(generator, sel):
        """extensions.py:54-75: the generator's own selection, or -- when it was left empty -- every varlev of its dataset."""
        if sel and 'varlev' in sel:
            return [str(v) for v in sel['varlev']]
        ds = getattr(generator, 'ds', None)
        if sel or ds is None or 'varlev' not in ds.dims:
            raise NotImplementedError('TimeSeriesEstimator needs a flat varlev dimension (variable / level selections are '
                                      'not mapped here)')
        return [str(v) for v in np.asarray(ds.coords['varlev'].values)]

    def _device_ok(self):
        m = self.model
        if getattr(m, 'impute', False) or getattr(m, 'scaler_type', None) is not None:
            return False
        if isinstance(m, DLWPFunctional) and m._n_steps > 1:
            return False
        return hasattr(m.model, 'engine') and not m.model._multi()

    def predict(self, steps, impute=False, keep_time_dim=False, prefer_first_times=True, **kwargs):
        """extensions.py:137-303, same arguments; returns the forecast with `f_hour` first."""
        if int(steps) < 1:
            raise ValueError('must use positive integer for steps')
        steps = int(steps)
        t_in, t_out = self._input_time_steps, self._output_time_steps
        if t_out <= t_in:
            keep_inputs, es = True, t_out
        else:
            keep_inputs = False
            es = t_in if prefer_first_times else t_out
        effective_steps = int(np.ceil(steps / es))
        gen = self.generator
        p, t = gen.generate([], scale_and_impute=False)
        p = np.ascontiguousarray(p, np.float32)
        t_shape = (t[0] if isinstance(t, (list, tuple)) else t).shape
        S = p.shape[0]
        H, W = gen.convolution_shape[-2:]
        in_vl, out_vl = list(self._input_sel['varlev']), list(self._output_sel['varlev'])
        shared = list(self._outputs_in_inputs['varlev'])
        idx_in, idx_out = [in_vl.index(v) for v in shared], [out_vl.index(v) for v in shared]
        V_in, V_out = len(in_vl), len(out_vl)
        times = np.asarray(gen.sample_times).astype('datetime64[s]')[:gen._n_sample]
        sample_coord = times.copy()
        shift = es + self._interval - 1
        lat, lon = np.asarray(gen.lat), np.asarray(gen.lon)

        if isinstance(self.model, DLWPFunctional) and self.model._n_steps > 1:      # extensions.py:204-208
            result = self.model.predict_timeseries(p, steps, keep_time_dim=True, **kwargs)
            result = result.reshape((-1,) + t_shape)[:effective_steps]
        elif self._device_ok():
            result = self._predict_device(p, effective_steps, S, t_in, t_out, V_in, V_out, H, W, shift, es, keep_inputs,
                                          prefer_first_times, impute, idx_in, idx_out, in_vl, times, lat, lon)
        else:
            result = self._predict_host(p, effective_steps, S, t_in, t_out, V_in, V_out, H, W, shift, es, keep_inputs,
                                        prefer_first_times, impute, idx_in, idx_out, in_vl, times, lat, lon, kwargs)
        result = np.asarray(result, np.float32).reshape((effective_steps, S, t_out, V_out, H, W))

        dt = self._dt
        time_coord = sample_coord + (t_in - 1) * dt
        if keep_time_dim:                                                           # extensions.py:259-273
            f_hour = np.arange(1, effective_steps * (es + self._interval - 1) + 1, es + self._interval - 1) * dt
            return _wrap(result, ['f_hour', 'time', 'time_step', 'varlev', 'lat', 'lon'],
                         [f_hour, time_coord, np.arange(t_out), np.array(out_vl), lat, lon])
        if not keep_inputs and prefer_first_times:                                  # extensions.py:276-278
            result = result[:, :, :es]
        result = result.transpose((0, 2, 1, 3, 4, 5))
        result = result.reshape((-1,) + result.shape[2:])
        f_hour = np.array([(np.arange(0, es) + self._interval + e * (es - 1 + self._interval)) * dt
                           for e in range(effective_steps)]).flatten()
        return _wrap(result[:steps], ['f_hour', 'time', 'varlev', 'lat', 'lon'],
                     [f_hour[:steps], time_coord, np.array(out_vl), lat, lon])

    # -- the reference's loop on the host (scalers / imputers / multi-device models) -----------------------------------------
    def _predict_host(self, p, effective_steps, S, t_in, t_out, V_in, V_out, H, W, shift, es, keep_inputs,
                      prefer_first_times, impute, idx_in, idx_out, in_vl, times, lat, lon, kwargs):
        p_shape = p.shape
        p = p.reshape(S, t_in, V_in, H, W).copy()
        p_mean = p.mean(axis=0) if impute else None
        result = np.full((effective_steps, S, t_out, V_out, H, W), np.nan, np.float32)
        for s in range(effective_steps):
            if kwargs.get('verbose', 0) > 0:
                print('Time step %d/%d' % (s + 1, effective_steps))
            r = np.asarray(self.model.predict(p.reshape(p_shape), **kwargs), np.float32).reshape(S, t_out, V_out, H, W)
            result[s] = r
            times = times + shift * self._dt
            q = np.full_like(p, np.nan)
            if shift < S:
                q[:S - shift] = p[shift:]
            p = q
            if impute:
                p[-es:] = p_mean[None]
            if self._add_insolation:
                k = in_vl.index('SOL')
                for n in range(t_in):
                    p[-es:, n, k] = insolation(times[-es:] + n * self._dt, lat, lon)
            if keep_inputs:
                p[:, t_in - es:, idx_in] = r[:, :, idx_out]
            elif prefer_first_times:
                p[:, :, idx_in] = r[:, :t_in][:, :, idx_out]
            else:
                p[:, :, idx_in] = r[:, t_out - t_in:][:, :, idx_out]
        return result

    # -- the same loop with the predictor array resident on the GPU ------------------------------------------------------------
    def _predict_device(self, p, effective_steps, S, t_in, t_out, V_in, V_out, H, W, shift, es, keep_inputs,
                        prefer_first_times, impute, idx_in, idx_out, in_vl, times, lat, lon):
        import torch
        eng = self.model.model.engine(S)
        if eng.max_batch < S:
            raise MemoryError('the %d samples of the generator do not fit one plan (capacity %d)' % (S, eng.max_batch))
        lib = nat.lib()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        hw = H * W
        cur = torch.from_numpy(p.reshape(S, t_in * V_in, H, W)).cuda()
        nxt = torch.empty_like(cur)
        mean = cur.mean(dim=0, keepdim=True) if impute else None
        series = torch.empty((effective_steps, S, t_out * V_out, H, W), dtype=torch.float32, device='cuda')
        # runs of consecutive (input, output) varlev positions: one strided copy per run and time step
        runs = []
        for a, b in zip(idx_in, idx_out):
            if runs and runs[-1][0] + runs[-1][2] == a and runs[-1][1] + runs[-1][2] == b:
                runs[-1][2] += 1
            else:
                runs.append([a, b, 1])
        if keep_inputs:
            pairs = [(t_in - es + k, k) for k in range(es)]          # (input time step, output time step)
        elif prefer_first_times:
            pairs = [(k, k) for k in range(t_in)]
        else:
            pairs = [(k, t_out - t_in + k) for k in range(t_in)]
        eng.lib.dlwp_debug_flags()
        # Which inputs are known (not NaN) is tracked on the host, per (sample, time step, varlev): the re-indexing runs
        # `shift` samples past the data per step, imputing / insolation / re-inserted outputs fill some of it back in.
        known = np.isfinite(p.reshape(S, t_in, V_in, -1)).all(axis=3)
        for s in range(effective_steps):
            # Only the valid samples go through the network: a NaN sample would poison the shared power-of-two exponent of
            # the tensor-core images (their amax); the reference's NaN rows come out as NaN here too.
            ok = known.all(axis=(1, 2))
            valid = int(np.argmin(ok)) if not ok.all() else S
            if ok[valid:].any():           # a hole in the sample axis (never produced by this loop): take the host path
                return self._predict_host(p, effective_steps, S, t_in, t_out, V_in, V_out, H, W, shift, es, keep_inputs,
                                          prefer_first_times, impute, idx_in, idx_out, in_vl,
                                          times - s * shift * self._dt, lat, lon, {})
            if valid > 0:
                eng.forward_into(cur[:valid], [series[s][:valid]])
            if valid < S:
                series[s][valid:].fill_(float('nan'))
            nk = np.zeros_like(known)
            if shift < S:
                nk[:S - shift] = known[shift:]
            if impute:
                nk[S - es:] = True
            if self._add_insolation:
                nk[S - es:, :, in_vl.index('SOL')] = True
            for ti, to in pairs:
                nk[:, ti, idx_in] = ok[:, None]
            known = nk
            times = times + shift * self._dt
            # p_da.reindex(sample=...) (extensions.py:226): sample i takes the predictors of sample i + shift
            nxt.fill_(float('nan'))
            if shift < S:
                nxt[:S - shift].copy_(cur[shift:])
            if impute:
                nxt[S - es:] = mean
            if self._add_insolation:                                  # extensions.py:234-238
                k = in_vl.index('SOL')
                sol = np.stack([insolation(times[-es:] + n * self._dt, lat, lon) for n in range(t_in)], axis=1)
                sol_d = torch.from_numpy(np.ascontiguousarray(sol)).cuda()          # (es, t_in, H, W)
                for n in range(t_in):
                    nxt[S - es:, n * V_in + k] = sol_d[:, n]
            # the outputs that are also inputs replace them (extensions.py:242-252)
            for ti, to in pairs:
                for a, b, n in runs:
                    src = series[s].data_ptr() + 4 * (to * V_out + b) * hw
                    dst = nxt.data_ptr() + 4 * (ti * V_in + a) * hw
                    nat.check(lib.dlwp_copy4d(src, dst, S, n, H, W, t_out * V_out * hw, hw, W, t_in * V_in * hw, hw, W,
                                              stream), 'dlwp_copy4d')
            cur, nxt = nxt, cur
        out = series.cpu().numpy()
        if eng.check_flags():      # non-finite data in VALID samples / precision underflow: the host loop (fp32 fallback inside)
            return self._predict_host(p, effective_steps, S, t_in, t_out, V_in, V_out, H, W, shift, es, keep_inputs,
                                      prefer_first_times, impute, idx_in, idx_out, in_vl,
                                      times - effective_steps * shift * self._dt, lat, lon, {})
        return out
