"""
Numpy restatement of the per-layer arithmetic on the rollout hot path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Every function works on channels_first arrays ``(N, C, H, W)`` unless ``data_format='channels_last'`` is passed, and
computes in the dtype of its input (tests call it with float64 = "tier-0" oracle).
"""

import numpy as np


# ---------------------------------------------------------------------------------------------------------------- #
# Argument normalisation (Keras ZeroPadding2D.__init__, inherited by PeriodicPadding2D at DLWP/custom.py:183-189)
# ---------------------------------------------------------------------------------------------------------------- #

def normalize_padding(padding):
    """int p -> ((p,p),(p,p)); (a,b) -> ((a,a),(b,b)); ((t,b),(l,r)) unchanged.  a = rows (lat), b = cols (lon)."""
    if isinstance(padding, (int, np.integer)):
        p = int(padding)
        return (p, p), (p, p)
    if len(padding) != 2:
        raise ValueError('`padding` should have two elements. Found: ' + str(padding))
    out = []
    for k, p in enumerate(padding):
        if isinstance(p, (int, np.integer)):
            out.append((int(p), int(p)))
        else:
            if len(p) != 2:
                raise ValueError('`padding[%d]` should be an int or a tuple of 2 ints. Found: %s' % (k, str(p)))
            out.append((int(p[0]), int(p[1])))
    return tuple(out)


def normalize_pair(v, name='value'):
    if isinstance(v, (int, np.integer)):
        return int(v), int(v)
    v = tuple(int(a) for a in v)
    if len(v) != 2:
        raise ValueError('`%s` should be an int or a tuple of 2 ints. Found: %s' % (name, str(v)))
    return v


def _axes(data_format):
    if data_format in (None, 'channels_first'):
        return 1, 2, 3  # c, h, w
    if data_format == 'channels_last':
        return 3, 1, 2
    raise ValueError("data_format must be 'channels_first' or 'channels_last'")


# ---------------------------------------------------------------------------------------------------------------- #
# Padding
# ---------------------------------------------------------------------------------------------------------------- #

def periodic_pad2d(x, padding, data_format='channels_first'):
    """
    DLWP/custom.py:191-214.  W is padded first with [x[..., W-l:W], x, x[..., 0:r]], THEN H of the already W-padded
    tensor with [o[H-t:H], o, o[0:b]] -- corners therefore wrap in both dimensions.  Python slice semantics are kept:
    a zero 'top' pad produces slice(H, H) (empty), which is what the reference relies on.
    """
    (t, b), (l, r) = normalize_padding(padding)
    _, ah, aw = _axes(data_format)
    H, W = x.shape[ah], x.shape[aw]

    def take(a, axis, sl):
        idx = [slice(None)] * a.ndim
        idx[axis] = sl
        return a[tuple(idx)]

    o = np.concatenate([take(x, aw, slice(W - l, W)), x, take(x, aw, slice(0, r))], axis=aw)
    o = np.concatenate([take(o, ah, slice(H - t, H)), o, take(o, ah, slice(0, b))], axis=ah)
    return o


def fill_pad2d(x, padding, data_format='channels_first'):
    """DLWP/custom.py:359-402 (FillPadding2D.call): the first / last row is stacked `pad` times above / below, then the
    first / last column of the row-padded tensor left / right -- np.pad(mode='edge') on H, then on W."""
    (t, b), (l, r) = normalize_padding(padding)
    _, ah, aw = _axes(data_format)
    pads = [(0, 0)] * x.ndim
    pads[ah] = (t, b)
    o = np.pad(x, pads, mode='edge')
    pads = [(0, 0)] * x.ndim
    pads[aw] = (l, r)
    return np.pad(o, pads, mode='edge')


def tf_pad2d(x, padding, mode='CONSTANT', data_format='channels_first'):
    """DLWP/custom.py:581-586 (TFPadding2D.call -> tf.pad): CONSTANT zeros, REFLECT (border excluded) or SYMMETRIC
    (border included) -- numpy's 'constant' / 'reflect' / 'symmetric' (tf.pad's documented semantics; tensorflow is not
    in /root/reference: unpinned)."""
    (t, b), (l, r) = normalize_padding(padding)
    _, ah, aw = _axes(data_format)
    pads = [(0, 0)] * x.ndim
    pads[ah] = (t, b)
    pads[aw] = (l, r)
    return np.pad(x, pads, mode={'CONSTANT': 'constant', 'REFLECT': 'reflect', 'SYMMETRIC': 'symmetric'}[mode.upper()])


def normalize_padding3d(padding):
    """keras ZeroPadding3D.__init__: int / 3 ints / 3 pairs -> ((a0,a1),(b0,b1),(c0,c1))."""
    if isinstance(padding, (int, np.integer)):
        return ((int(padding),) * 2,) * 3
    if len(padding) != 3:
        raise ValueError('`padding` should have 3 elements. Found: ' + str(padding))
    return tuple(normalize_pair(p, 'entry of padding') for p in padding)


def periodic_pad3d(x, padding, data_format='channels_first'):
    """
    DLWP/custom.py:277-306 on a 5-D array.  channels_first: the padded axes are 2, 3, 4 -- in the recurrent example nets
    the input is (batch, time, channels, lat, lon), so 'depth' is time and the first padded axis is the channel axis
    (examples/train.py:144-149).  Order: axis 4 (horizontal), then axis 3 of the result, then axis 2 of that result.
    """
    (d0, d1), (t, b), (l, r) = normalize_padding3d(padding)
    a1, a2, a3 = (2, 3, 4) if data_format in (None, 'channels_first') else (1, 2, 3)

    def take(a, axis, sl):
        idx = [slice(None)] * a.ndim
        idx[axis] = sl
        return a[tuple(idx)]

    n1, n2, n3 = x.shape[a1], x.shape[a2], x.shape[a3]
    o = np.concatenate([take(x, a3, slice(n3 - l, n3)), x, take(x, a3, slice(0, r))], axis=a3)
    o = np.concatenate([take(o, a2, slice(n2 - t, n2)), o, take(o, a2, slice(0, b))], axis=a2)
    o = np.concatenate([take(o, a1, slice(n1 - d0, n1)), o, take(o, a1, slice(0, d1))], axis=a1)
    return o


def zero_pad3d(x, padding, data_format='channels_first'):
    """Keras ZeroPadding3D: constant zeros on the three padded axes."""
    pd = normalize_padding3d(padding)
    axes = (2, 3, 4) if data_format in (None, 'channels_first') else (1, 2, 3)
    pads = [(0, 0)] * x.ndim
    for a, p in zip(axes, pd):
        pads[a] = p
    return np.pad(x, pads, mode='constant')


def hard_sigmoid(v):
    """keras.backend.hard_sigmoid: clip(0.2 x + 0.5, 0, 1)."""
    return np.clip(0.2 * v + 0.5, 0.0, 1.0)


def conv_lstm2d(x, kernel, recurrent_kernel, bias, dilation=(1, 1), padding='valid', act='tanh',
                recurrent_act='hard_sigmoid', return_sequences=True):
    """
    keras.layers.ConvLSTM2D forward (Keras 2.2 ConvLSTM2DCell.call; Keras itself is not in /root/reference, so this is
    a restatement of its published cell: PARITY UNPINNED), channels_first: x (N, T, C, H, W); kernel (kh, kw, C, 4F),
    recurrent_kernel (kh, kw, F, 4F), bias (4F,) or None; gate blocks i, f, c, o.  Initial h and c are zero.
        x_g = conv(x_t, W_g, padding, dilation) + b_g          (the layer's padding and dilation)
        h_g = conv(h_{t-1}, U_g, 'same', no dilation)          (recurrent_conv)
        i = s(x_i + h_i); f = s(x_f + h_f); c = f * c_{t-1} + i * a(x_c + h_c); o = s(x_o + h_o); h = o * a(c)
    """
    n, T, C, H, W = x.shape
    kh, kw, _, F4 = kernel.shape
    F = F4 // 4
    dh, dw = normalize_pair(dilation, 'dilation_rate')
    a = activation(act)
    s = hard_sigmoid if recurrent_act == 'hard_sigmoid' else (lambda v: 1.0 / (1.0 + np.exp(-v)))
    if padding == 'same':
        th, tw = dh * (kh - 1), dw * (kw - 1)
        x = np.pad(x, [(0, 0), (0, 0), (0, 0), (th // 2, th - th // 2), (tw // 2, tw - tw // 2)], mode='constant')
    rp = [(0, 0), (0, 0), ((kh - 1) // 2, kh - 1 - (kh - 1) // 2), ((kw - 1) // 2, kw - 1 - (kw - 1) // 2)]
    h = c = None
    out = []
    for t in range(T):
        z = conv2d_valid(x[:, t], kernel, bias, (dh, dw))
        if h is not None:
            z = z + conv2d_valid(np.pad(h, rp, mode='constant'), recurrent_kernel, None, (1, 1))
        zi, zf, zc, zo = z[:, :F], z[:, F:2 * F], z[:, 2 * F:3 * F], z[:, 3 * F:]
        i, f, o = s(zi), s(zf), s(zo)
        c = i * a(zc) if c is None else f * c + i * a(zc)
        h = o * a(c)
        out.append(h)
    return np.stack(out, axis=1) if return_sequences else h


def zero_pad2d(x, padding, data_format='channels_first'):
    """Keras ZeroPadding2D: constant-zero rows/cols."""
    (t, b), (l, r) = normalize_padding(padding)
    _, ah, aw = _axes(data_format)
    pads = [(0, 0)] * x.ndim
    pads[ah] = (t, b)
    pads[aw] = (l, r)
    return np.pad(x, pads, mode='constant')


# ---------------------------------------------------------------------------------------------------------------- #
# Activations
# ---------------------------------------------------------------------------------------------------------------- #

def activation(name):
    if name is None or name == 'linear':
        return lambda v: v
    if name == 'tanh':
        return np.tanh
    if name == 'relu':
        return lambda v: np.maximum(v, 0)
    if callable(name):
        return name
    raise ValueError('unsupported activation %r' % (name,))


# ---------------------------------------------------------------------------------------------------------------- #
# Conv2D ('valid', stride 1, dilation d): Keras semantics, SURVEY.md Appendix A.2
# ---------------------------------------------------------------------------------------------------------------- #

def conv2d_valid(x, kernel, bias=None, dilation=(1, 1), strides=(1, 1), data_format='channels_first'):
    """
    Cross-correlation (no kernel flip): out[n,o,y,x] = b[o] + sum_{c,i,j} k[i,j,c,o] * in[n,c, y*s+d*i, x*s+d*j].
    ``kernel`` is in Keras layout (kh, kw, Cin, Cout).  Direct summation tap by tap (no FFT / no Winograd) so that a
    float64 call is an exact-to-rounding reference.
    """
    if data_format == 'channels_last':
        y = conv2d_valid(np.moveaxis(x, 3, 1), kernel, bias, dilation, strides, 'channels_first')
        return np.moveaxis(y, 1, 3)
    dh, dw = normalize_pair(dilation, 'dilation_rate')
    sh, sw = normalize_pair(strides, 'strides')
    kh, kw, cin, cout = kernel.shape
    n, c, H, W = x.shape
    if c != cin:
        raise ValueError('input has %d channels but kernel expects %d' % (c, cin))
    Ho = (H - dh * (kh - 1) - 1) // sh + 1
    Wo = (W - dw * (kw - 1) - 1) // sw + 1
    if Ho <= 0 or Wo <= 0:
        raise ValueError('negative output size')
    out = np.zeros((n, cout, Ho, Wo), dtype=np.result_type(x.dtype, kernel.dtype))
    for i in range(kh):
        for j in range(kw):
            patch = x[:, :, i * dh: i * dh + (Ho - 1) * sh + 1: sh, j * dw: j * dw + (Wo - 1) * sw + 1: sw]
            out += np.einsum('nchw,co->nohw', patch, kernel[i, j], optimize=True)
    if bias is not None:
        out += np.asarray(bias).reshape(1, cout, 1, 1)
    return out


def pad_conv2d_closed_form(x, kernel, bias, dilation, pad_h, pad_w, mode_h, mode_w):
    """
    The fused form the CUDA kernels implement (SURVEY.md 8c "Oracle definition"):
        out[n,o,y,x] = b[o] + sum k[i,j,c,o] * X(n, c, y + d*i - pad_t, x + d*j - pad_l)
    with X wrapped modulo the axis length for mode 'periodic' and 0 outside the axis for mode 'zero'.  Independent of
    periodic_pad2d/zero_pad2d above (index arithmetic instead of concatenation) so the two can be checked against each
    other.
    """
    dh, dw = normalize_pair(dilation)
    (pt, pb), (pl, pr) = pad_h, pad_w
    kh, kw, cin, cout = kernel.shape
    n, c, H, W = x.shape
    Ho = H + pt + pb - dh * (kh - 1)
    Wo = W + pl + pr - dw * (kw - 1)
    out = np.zeros((n, cout, Ho, Wo), dtype=np.result_type(x.dtype, kernel.dtype))
    ys = np.arange(Ho)
    xs = np.arange(Wo)
    for i in range(kh):
        yy = ys + dh * i - pt
        if mode_h == 'periodic':
            vy = np.ones(Ho, bool)
            yy = np.mod(yy, H)
        else:
            vy = (yy >= 0) & (yy < H)
            yy = np.clip(yy, 0, H - 1)
        for j in range(kw):
            xx = xs + dw * j - pl
            if mode_w == 'periodic':
                vx = np.ones(Wo, bool)
                xx = np.mod(xx, W)
            else:
                vx = (xx >= 0) & (xx < W)
                xx = np.clip(xx, 0, W - 1)
            patch = x[:, :, yy][:, :, :, xx] * (vy[:, None] & vx[None, :])
            out += np.einsum('nchw,co->nohw', patch, kernel[i, j], optimize=True)
    if bias is not None:
        out += np.asarray(bias).reshape(1, cout, 1, 1)
    return out


# ---------------------------------------------------------------------------------------------------------------- #
# Pool / upsample / slice / concat
# ---------------------------------------------------------------------------------------------------------------- #

def max_pool2d(x, pool_size=2, data_format='channels_first'):
    """Keras MaxPooling2D(pool_size) with default strides=pool_size and padding='valid' (floor)."""
    ph, pw = normalize_pair(pool_size, 'pool_size')
    if data_format == 'channels_last':
        return np.moveaxis(max_pool2d(np.moveaxis(x, 3, 1), pool_size), 1, 3)
    n, c, H, W = x.shape
    Ho, Wo = H // ph, W // pw
    v = x[:, :, :Ho * ph, :Wo * pw].reshape(n, c, Ho, ph, Wo, pw)
    return v.max(axis=(3, 5))


def upsample2d(x, size=2, data_format='channels_first'):
    """Keras UpSampling2D(size): nearest-neighbour repeat on H then W."""
    sh, sw = normalize_pair(size, 'size')
    _, ah, aw = _axes(data_format)
    return np.repeat(np.repeat(x, sh, axis=ah), sw, axis=aw)


def slice_channels(x, start, end, step=None, axis=1):
    """DLWP/custom.py:675-692 (slice_layer): x[:, start:end:step] along ``axis``; axis < 0 is rejected."""
    if axis < 0:
        raise ValueError("'slice_layer' can only work on a specified axis > 0")
    idx = [slice(None)] * axis + [slice(start, end, step)]
    return x[tuple(idx)]


# ---------------------------------------------------------------------------------------------------------------- #
# RowConnected2D (DLWP/custom.py:782-837, 840-896)
# ---------------------------------------------------------------------------------------------------------------- #

def row_conv2d(x, kernel, bias=None, strides=(1, 1), data_format='channels_first'):
    """
    ``kernel``: (H_out, kh, kw, Cin, Cout); ``bias``: (H_out, 1, Cout).  One 'valid' conv per output row over the slab
    rows [i*stride_row, i*stride_col + kh) (custom.py:881 -- well defined only for equal strides; (1,1) is the only
    use), results concatenated along H.  Bias is added per (row, filter): K.bias_add with a (H_out, 1, Cout) bias
    broadcasts over W (custom.py:834).
    """
    if data_format == 'channels_last':
        return np.moveaxis(row_conv2d(np.moveaxis(x, 3, 1), kernel, bias, strides), 1, 3)
    sr, sc = normalize_pair(strides, 'strides')
    Ho, kh = kernel.shape[0], kernel.shape[1]
    rows = []
    for i in range(Ho):
        slab = x[:, :, i * sr: i * sc + kh, :]
        rows.append(conv2d_valid(slab, kernel[i], None, (1, 1), (sr, sc)))
    out = np.concatenate(rows, axis=2)
    if bias is not None:
        # (H_out, 1, Cout) -> (1, Cout, H_out, 1)
        out = out + np.transpose(np.asarray(bias), (2, 0, 1))[None]
    return out


# ---------------------------------------------------------------------------------------------------------------- #
# Weight initialisation used by every synthetic benchmark / test
# ---------------------------------------------------------------------------------------------------------------- #

def glorot_uniform(rng, kh, kw, cin, cout, dtype=np.float32):
    """Keras default kernel initialiser: U(-L, L), L = sqrt(6 / (fan_in + fan_out)), drawn in (kh,kw,Cin,Cout) order."""
    fan_in, fan_out = kh * kw * cin, kh * kw * cout
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=(kh, kw, cin, cout)).astype(dtype)
