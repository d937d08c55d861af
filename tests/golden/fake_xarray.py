"""
A minimal stand-in for the slice of xarray that `TimeSeriesEstimator.predict` (DLWP/model/extensions.py:136-303) touches,
so that the REFERENCE's own loop can be executed here (xarray / netCDF4 are not installed) to produce golden vectors for
`oracle/estimator.py` and `dlwp_b200/model/extensions.py` (tests/golden/make_golden.py:gen_estimator).  TEST INFRASTRUCTURE.

Semantics implemented (xarray's documented behaviour for these calls):
* `DataArray(data, coords=[...], dims=[...])`: one 1-d coordinate per dimension; `.values`, `.shape`, `.dims`, attribute and
  item access to coordinates (`da.sample`, `da['sample']`), arithmetic with scalars / 0-d arrays (element-wise on the values,
  coordinates kept).
* positional `da[idx]` returns a VIEW of the values for basic indexing (so `da.loc[{'varlev': 'SOL'}][-es:] = x` writes through,
  as it does in xarray); `da[idx] = v` assigns positionally.
* `da.loc[{dim: labels}]`: labels -> positions by exact match on the coordinate (scalar label: the dimension is dropped, the
  result is a view; array of labels: fancy indexing, a copy); `da.loc[{...}] = value` writes the selected block, a DataArray
  value being assigned by position after a check that the shapes agree.
* `da.reindex(sample=new, method=None)`: rows are looked up by exact label; labels absent from the old index give NaN rows.
* `da.isel(dim=slice)`, `da.sel(dim=labels)` (= `.loc`), `da.load()` (no-op), `Dataset` with coordinate and data variables.
"""

import numpy as np


def _raw(x):
    return x.values if isinstance(x, DataArray) else x


class _Loc(object):
    def __init__(self, da):
        self.da = da

    def _index(self, sel):
        idx, drop = [slice(None)] * self.da.values.ndim, []
        for dim, labels in sel.items():
            axis = self.da.dims.index(dim)
            coord = np.asarray(self.da.coords[dim])
            lab = np.asarray(_raw(labels))
            if lab.ndim == 0:
                pos = np.nonzero(coord == lab)[0]
                if len(pos) != 1:
                    raise KeyError(labels)
                idx[axis] = int(pos[0])
                drop.append(dim)
            else:
                where = []
                for l in lab:
                    pos = np.nonzero(coord == l)[0]
                    if len(pos) != 1:
                        raise KeyError(l)
                    where.append(int(pos[0]))
                idx[axis] = np.array(where, dtype=np.int64)
        return idx, drop

    def __getitem__(self, sel):
        idx, drop = self._index(sel)
        da = self.da
        values = da.values[tuple(i if isinstance(i, int) else slice(None) for i in idx)]   # scalar labels: a view
        dims = [d for d in da.dims if d not in drop]
        coords = {d: da.coords[d] for d in dims}
        for axis, d in enumerate(dims):                   # label arrays: orthogonal indexing, one axis at a time (copies)
            ix = idx[da.dims.index(d)]
            if isinstance(ix, np.ndarray):
                values = np.take(values, ix, axis=axis)
                coords[d] = coords[d][ix]
        return DataArray(values, coords=[coords[d] for d in dims], dims=dims, _share=True)

    def __setitem__(self, sel, value):
        idx, _ = self._index(sel)
        view = self.da.values[tuple(i if isinstance(i, int) else slice(None) for i in idx)]
        kept = [i for i in idx if not isinstance(i, int)]
        rest = [ix if isinstance(ix, np.ndarray) else np.arange(view.shape[a]) for a, ix in enumerate(kept)]
        v = np.asarray(_raw(value))
        if v.ndim and v.shape != tuple(len(r) for r in rest):
            raise ValueError('shape mismatch in .loc assignment: %r vs %r' % (v.shape, tuple(len(r) for r in rest)))
        view[np.ix_(*rest)] = v                           # writes through to the array


class DataArray(object):
    def __init__(self, data, coords=None, dims=None, _share=False):
        self.values = data if _share else np.array(_raw(data))
        if dims is None:
            dims = ['dim_%d' % i for i in range(self.values.ndim)]
        self.dims = tuple(dims)
        coords = [] if coords is None else coords
        if isinstance(coords, dict):                      # xr.DataArray(data, coords={'sample': ..., ...}, dims=[...])
            coords = [coords[d] for d in self.dims]
        if self.values.ndim and len(coords) != self.values.ndim:
            raise ValueError('one coordinate per dimension expected')
        self.coords = {}
        for d, c, n in zip(self.dims, coords, self.values.shape):
            c = np.asarray(_raw(c)) if not isinstance(c, range) else np.arange(c.start, c.stop, c.step)
            if c.shape != (n,):
                raise ValueError('coordinate %r has shape %r, dimension has %d entries' % (d, c.shape, n))
            self.coords[d] = c

    # -- basics --------------------------------------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.values.shape

    def __len__(self):
        return len(self.values)

    def __iter__(self):
        return iter(self.values)

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def _coord(self, name):
        return DataArray(self.coords[name], coords=[self.coords[name]], dims=[name])

    def __getattr__(self, name):
        coords = self.__dict__.get('coords', {})
        if name in coords:
            return self._coord(name)
        raise AttributeError(name)

    @property
    def loc(self):
        return _Loc(self)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._coord(key)
        key = key if isinstance(key, tuple) else (key,)
        key = tuple(_raw(k) for k in key)
        values = self.values[key]
        dims, coords = [], []
        for axis, d in enumerate(self.dims):
            k = key[axis] if axis < len(key) else slice(None)
            if isinstance(k, (int, np.integer)):
                continue
            dims.append(d)
            coords.append(self.coords[d][k])
        return DataArray(values, coords=coords, dims=dims, _share=True)

    def __setitem__(self, key, value):
        key = key if isinstance(key, tuple) else (key,)
        self.values[tuple(_raw(k) for k in key)] = np.asarray(_raw(value))

    # -- arithmetic ------------------------------------------------------------------------------------------------------------
    def _binary(self, other, op):
        out = op(self.values, _raw(other))
        if np.ndim(out) == 0:
            return DataArray(out)
        return DataArray(out, coords=[self.coords[d] for d in self.dims], dims=self.dims)

    def __add__(self, o):
        return self._binary(o, lambda a, b: a + b)

    def __radd__(self, o):
        return self._binary(o, lambda a, b: b + a)

    def __sub__(self, o):
        return self._binary(o, lambda a, b: a - b)

    def __mul__(self, o):
        return self._binary(o, lambda a, b: a * b)

    def __rmul__(self, o):
        return self._binary(o, lambda a, b: b * a)

    # -- label operations --------------------------------------------------------------------------------------------------------
    def reindex(self, method=None, **indexers):
        assert method is None and len(indexers) == 1
        (dim, new), = indexers.items()
        new = np.asarray(_raw(new))
        axis = self.dims.index(dim)
        old = self.coords[dim]
        shape = list(self.values.shape)
        shape[axis] = len(new)
        out = np.full(shape, np.nan, dtype=self.values.dtype)
        for i, label in enumerate(new):
            pos = np.nonzero(old == label)[0]
            if len(pos):
                src, dst = [slice(None)] * out.ndim, [slice(None)] * out.ndim
                src[axis], dst[axis] = int(pos[0]), i
                out[tuple(dst)] = self.values[tuple(src)]
        coords = [new if d == dim else self.coords[d] for d in self.dims]
        return DataArray(out, coords=coords, dims=self.dims)

    def sel(self, **indexers):
        return self.loc[indexers] if indexers else self

    def load(self):
        return self

    def isel(self, **indexers):
        key = tuple(indexers.get(d, slice(None)) for d in self.dims)
        return self[key]


class Dataset(object):
    """Just the attributes of `generator.ds` the estimator reads: dims, variables, coords, item / attribute access."""

    def __init__(self, coords, **data_vars):
        self.coords = {k: DataArray(v, coords=[v], dims=[k]) for k, v in coords.items()}
        self.dims = {k: len(v) for k, v in coords.items()}
        self.variables = dict(self.coords)
        for name, da in data_vars.items():                # e.g. predictors=DataArray(...)
            self.variables[name] = da
            self.__dict__[name] = da

    def load(self):
        return self

    def __getitem__(self, name):
        return self.coords[name]

    def __getattr__(self, name):
        coords = self.__dict__.get('coords', {})
        if name in coords:
            return coords[name]
        raise AttributeError(name)
