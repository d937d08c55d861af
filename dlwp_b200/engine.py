"""
Lowering of a (front-end) Keras graph to a libdlwp_b200 plan, and the object that runs it.

This is what stands where `keras.Model.predict` -> `tf.Session.run` stands in the reference
(DLWP/model/models.py:241, :412): the layer DAG recorded by dlwp_b200.keras is lowered ONCE to a static list of device ops
over (N, C, H, W) buffers:

* PeriodicPadding2D / ZeroPadding2D (DLWP/custom.py:191-214) become padding *attributes* of the following Conv2D -- no
  padded tensor exists in memory;
* `slice_layer` (custom.py:675-692) and `concatenate(axis=1)` become channel windows of shared buffers (a producer writes
  straight into its slot of the concatenated buffer);
* the last op of every model output writes directly into caller memory (for a rollout: into its slot of the series).

`CompiledNet` owns the DlwpPlan handle.  All compute goes through the C ABI; torch is used for device memory, pinned
host memory and the current stream only.
"""

import ctypes

import sys

import numpy as np

from . import _native as nat
from .keras import layers as KL
from .keras.engine import InputLayer


class _Val(object):
    """Symbolic value during lowering: a channel window of a buffer plus not-yet-applied paddings."""
    __slots__ = ('buf', 'c0', 'C', 'H', 'W', 'pads', 'shape')

    def __init__(self, buf, c0, C, H, W, pads=(), shape=None):
        self.buf, self.c0, self.C, self.H, self.W = buf, c0, C, H, W
        self.pads = tuple(pads)
        self.shape = shape if shape is not None else (C, H, W)  # logical per-sample shape reported to the caller


def _unsupported(layer, why):
    raise NotImplementedError('layer %s (%s) cannot be lowered to the GPU plan: %s' %
                              (layer.name, layer.__class__.__name__, why))


# FillPadding2D / TFPadding2D(REFLECT | SYMMETRIC) (DLWP/custom.py:309-402, 527-599): stand-alone pad ops, never fused
_STANDALONE_PAD_MODES = {'fill': nat.PAD_EDGE, 'reflect': nat.PAD_REFLECT, 'symmetric': nat.PAD_SYMMETRIC}


class Lowering(object):
    def __init__(self, model):
        self.model = model
        self.buffers = []   # dicts: kind, C, H, W, output_index
        self.ops = []       # dicts
        self.weight_layers = []
        self.out_vals = []
        self._lower()

    # -- helpers -------------------------------------------------------------------------------------------------
    def _new_buffer(self, C, H, W, kind=nat.BUF_INTERNAL, output_index=0):
        self.buffers.append({'kind': kind, 'C': C, 'H': H, 'W': W, 'output_index': output_index})
        return len(self.buffers) - 1

    def _weight_id(self, layer):
        if layer not in self.weight_layers:
            self.weight_layers.append(layer)
        return self.weight_layers.index(layer)

    def _op(self, kind, src, dst_buf, dst_c0=0, **kw):
        op = dict(kind=kind, src=src.buf, src_c0=src.c0, src_c=src.C, dst=dst_buf, dst_c0=dst_c0, weight_id=-1,
                  pad_t=0, pad_b=0, pad_l=0, pad_r=0, pad_mode_h=0, pad_mode_w=0, Cout=0, kh=0, kw=0, dil_h=1,
                  dil_w=1, act=0, pre_op=0, rowwise=0, impl=0, row_begin=0, row_end=0, aux=-1, aux_c0=0, act2=0)
        op.update(kw)
        self.ops.append(op)
        return op

    @staticmethod
    def _pad_mode(mode, layer):
        if mode == 'zero':
            return nat.PAD_ZERO
        if mode == 'periodic':
            return nat.PAD_PERIODIC
        if mode in _STANDALONE_PAD_MODES:
            return _STANDALONE_PAD_MODES[mode]
        _unsupported(layer, 'padding mode %r has no kernel' % mode)

    def _materialise_one_pad(self, v):
        """Apply the first pending padding with a stand-alone pad op."""
        mode, (t, b), (l, r), layer = v.pads[0]
        m = self._pad_mode(mode, layer)
        H, W = v.H + t + b, v.W + l + r
        buf = self._new_buffer(v.C, H, W)
        self._op(nat.OP_PAD, _Val(v.buf, v.c0, v.C, v.H, v.W), buf, pad_t=t, pad_b=b, pad_l=l, pad_r=r,
                 pad_mode_h=m, pad_mode_w=m)
        return _Val(buf, 0, v.C, H, W, v.pads[1:])

    def _materialise(self, v):
        while v.pads:
            v = self._materialise_one_pad(v)
        return v

    def _check_format(self, layer):
        """Every layer of a model uses the model's data_format (channels_last models compute channels_first between one
        layout op at each end: DLWP/custom.py:205-213 is the channels_last branch of PeriodicPadding2D)."""
        if layer.data_format != self.fmt:
            _unsupported(layer, 'data_format %r in a %s model' % (layer.data_format, self.fmt))

    def _settle_pad(self, v, layer):
        """Zero / periodic paddings stay pending (the next conv fuses them); the other modes become pad ops now."""
        if layer.pad_mode in ('zero', 'periodic'):
            return v
        self._pad_mode(layer.pad_mode, layer)      # raises for modes without a kernel
        shape = v.shape
        v = self._materialise(v)
        v.shape = shape
        return v

    def _fusable_pads(self, v):
        """Reduce pending paddings until at most one layer pads each axis; return (v, (mode_h,t,b), (mode_w,l,r))."""
        while True:
            h = [(p[0], p[1], p[3]) for p in v.pads if p[1] != (0, 0)]
            w = [(p[0], p[2], p[3]) for p in v.pads if p[2] != (0, 0)]
            if len(h) <= 1 and len(w) <= 1:
                hm = (self._pad_mode(h[0][0], h[0][2]),) + h[0][1] if h else (nat.PAD_ZERO, 0, 0)
                wm = (self._pad_mode(w[0][0], w[0][2]),) + w[0][1] if w else (nat.PAD_ZERO, 0, 0)
                return v, hm, wm
            v = self._materialise_one_pad(v)

    # -- per-layer emitters --------------------------------------------------------------------------------------
    def _emit(self, layer, ins):
        if isinstance(layer, KL.ZeroPadding3D):  # includes PeriodicPadding3D (DLWP/custom.py:217-306)
            # (batch, time, channels, lat, lon) declared channels_first: padding axes are (channels, lat, lon)
            # (examples/train.py:144-149).  Padding the two trailing axes is a 2-D padding of every (time, channel) plane.
            v = ins[0]
            if layer.data_format != 'channels_first' or len(v.shape) != 4:
                _unsupported(layer, "only 5-D channels_first inputs (batch, time, channels, lat, lon) are lowered")
            (c0, c1), (t, b), (l, r) = layer.padding
            if (c0, c1) != (0, 0):
                _unsupported(layer, 'padding of the first padded axis (the channel axis of a (time, channels, lat, lon) '
                                    'input) is not on the hot path')
            if layer.pad_mode == 'periodic' and (max(t, b) > v.shape[2] or max(l, r) > v.shape[3]):
                raise ValueError('PeriodicPadding3D %s pads by more than the axis length' % layer.name)
            pads = v.pads + ((layer.pad_mode, (t, b), (l, r), layer),)
            out = _Val(v.buf, v.c0, v.C, v.H, v.W, pads, tuple(v.shape[:2]) + (v.shape[2] + t + b, v.shape[3] + l + r))
            return self._settle_pad(out, layer)
        if isinstance(layer, KL.ConvLSTM2D):
            return self._emit_convlstm(layer, ins[0])
        if isinstance(layer, KL.ZeroPadding2D):  # includes PeriodicPadding2D & friends (pad_mode attribute)
            v = ins[0]
            self._check_format(layer)
            (t, b), (l, r) = layer.padding
            if layer.pad_mode == 'periodic' and (max(t, b) > v.H + sum(p[1][0] + p[1][1] for p in v.pads) or
                                                 max(l, r) > v.W + sum(p[2][0] + p[2][1] for p in v.pads)):
                raise ValueError('PeriodicPadding2D %s pads by more than the axis length' % layer.name)
            pads = v.pads + ((layer.pad_mode, (t, b), (l, r), layer),)
            out = _Val(v.buf, v.c0, v.C, v.H, v.W, pads, (v.C, v.H + sum(p[1][0] + p[1][1] for p in pads),
                                                          v.W + sum(p[2][0] + p[2][1] for p in pads)))
            return self._settle_pad(out, layer)
        if isinstance(layer, KL.Conv2D):  # includes RowConnected2D
            return self._emit_conv(layer, ins[0])
        if isinstance(layer, KL.MaxPooling2D):
            if layer.pool_size != (2, 2) or layer.strides != (2, 2) or layer.padding != 'valid':
                _unsupported(layer, 'only MaxPooling2D(2) with default strides / valid padding is implemented')
            self._check_format(layer)
            v = self._materialise(ins[0])
            buf = self._new_buffer(v.C, v.H // 2, v.W // 2)
            self._op(nat.OP_MAXPOOL, v, buf)
            return _Val(buf, 0, v.C, v.H // 2, v.W // 2)
        if isinstance(layer, KL.UpSampling2D):
            if layer.size != (2, 2):
                _unsupported(layer, 'only UpSampling2D(2) is implemented')
            self._check_format(layer)
            v = self._materialise(ins[0])
            buf = self._new_buffer(v.C, v.H * 2, v.W * 2)
            self._op(nat.OP_UPSAMPLE, v, buf)
            return _Val(buf, 0, v.C, v.H * 2, v.W * 2)
        if isinstance(layer, KL.ChannelSlice):
            v = ins[0]
            if layer.axis != (3 if self.fmt == 'channels_last' else 1) or layer.step not in (None, 1):
                _unsupported(layer, 'only contiguous channel slices (step 1 along the channel axis) are lowered')
            a, b, _ = slice(layer.start, layer.end, None).indices(v.C)
            if b <= a:
                raise ValueError('slice_layer %s selects no channels' % layer.name)
            return _Val(v.buf, v.c0 + a, b - a, v.H, v.W, v.pads, (b - a,) + tuple(v.shape[1:]))
        if isinstance(layer, KL.Concatenate):
            if layer.axis not in ((3, -1) if self.fmt == 'channels_last' else (1, -3)):
                _unsupported(layer, 'only concatenation along the channel axis is lowered')
            vs = [self._materialise(v) for v in ins]
            C = sum(v.C for v in vs)
            buf = self._new_buffer(C, vs[0].H, vs[0].W)
            c0 = 0
            for v in vs:
                self._op(nat.OP_COPY, v, buf, dst_c0=c0)
                c0 += v.C
            return _Val(buf, 0, C, vs[0].H, vs[0].W)
        if isinstance(layer, KL.Reshape):
            v = ins[0]
            if self.fmt == 'channels_last':
                _unsupported(layer, 'Reshape in a channels_last model')
            tgt = layer.compute_output_shape((None,) + tuple(v.shape))[1:]
            if len(tgt) >= 2 and tuple(tgt[-2:]) == tuple(v.shape[-2:]):
                return _Val(v.buf, v.c0, v.C, v.H, v.W, v.pads, tuple(tgt))  # regroups (T, C) <-> (T*C): a view
            _unsupported(layer, 'Reshape that changes the spatial axes')
        if isinstance(layer, KL.Lambda):
            _unsupported(layer, 'arbitrary Lambda functions run on the host; use DLWP.custom.slice_layer')
        _unsupported(layer, 'no kernel for this layer type')

    def _emit_conv(self, layer, v):
        if len(v.shape) != 3:
            _unsupported(layer, 'expects a 4-D (batch, channels, lat, lon) input, got per-sample shape %s' % (v.shape,))
        self._check_format(layer)
        if layer.strides != (1, 1):
            _unsupported(layer, 'strided convolutions are not on the hot path')
        act = nat.ACTIVATIONS.get(layer.activation)
        if act is None and layer.activation is not None:
            _unsupported(layer, 'activation %r' % (layer.activation,))
        kh, kw = layer.kernel_size
        dh, dw = layer.dilation_rate
        if layer.padding == 'same':
            th, tw = dh * (kh - 1), dw * (kw - 1)
            v = _Val(v.buf, v.c0, v.C, v.H, v.W,
                     v.pads + (('zero', (th // 2, th - th // 2), (tw // 2, tw - tw // 2), layer),))
        v, (mh, t, b), (mw, l, r) = self._fusable_pads(v)
        Ho = v.H + t + b - dh * (kh - 1)
        Wo = v.W + l + r - dw * (kw - 1)
        if Ho <= 0 or Wo <= 0:
            raise ValueError('Negative dimension size caused by the convolution of layer %s' % layer.name)
        rowwise = 1 if layer.__class__.__name__ == 'RowConnected2D' else 0
        buf = self._new_buffer(layer.filters, Ho, Wo)
        self._op(nat.OP_CONV, _Val(v.buf, v.c0, v.C, v.H, v.W), buf, weight_id=self._weight_id(layer), pad_t=t,
                 pad_b=b, pad_l=l, pad_r=r, pad_mode_h=mh, pad_mode_w=mw, Cout=layer.filters, kh=kh, kw=kw, dil_h=dh,
                 dil_w=dw, act=act or 0, rowwise=rowwise)
        return _Val(buf, 0, layer.filters, Ho, Wo)

    def _emit_convlstm(self, layer, v):
        """keras ConvLSTM2D on a (time, channels, lat, lon) value: per time step the input convolution (the layer's
        padding / dilation, pending periodic / zero paddings fused), the recurrent convolution of h_{t-1} ('same', zero
        padded, undilated -- keras ConvLSTM2DCell.recurrent_conv) and one gate op (dlwp_convlstm_gates).  h_t is written
        straight into its slot of the (time * filters) output, which the next step's recurrent convolution reads."""
        if layer.data_format != 'channels_first' or len(v.shape) != 4:
            _unsupported(layer, 'only 5-D channels_first inputs (batch, time, channels, lat, lon) are lowered')
        if layer.strides != (1, 1) or layer.go_backwards or layer.stateful:
            _unsupported(layer, 'strides / go_backwards / stateful are not on the hot path')
        act = nat.ACTIVATIONS.get(layer.activation)
        ract = nat.RECURRENT_ACTIVATIONS.get(layer.recurrent_activation)
        if act is None or ract is None:
            _unsupported(layer, 'activation %r / recurrent_activation %r' % (layer.activation, layer.recurrent_activation))
        T, C = v.shape[0], v.shape[1]
        F = layer.filters
        kh, kw = layer.kernel_size
        dh, dw = layer.dilation_rate
        if layer.padding == 'same':
            th, tw = dh * (kh - 1), dw * (kw - 1)
            v = _Val(v.buf, v.c0, v.C, v.H, v.W,
                     v.pads + (('zero', (th // 2, th - th // 2), (tw // 2, tw - tw // 2), layer),), v.shape)
        v, (mh, t, b), (mw, l, r) = self._fusable_pads(v)
        Ho = v.H + t + b - dh * (kh - 1)
        Wo = v.W + l + r - dw * (kw - 1)
        if Ho <= 0 or Wo <= 0:
            raise ValueError('Negative dimension size caused by the convolution of layer %s' % layer.name)
        z = self._new_buffer(8 * F, Ho, Wo)       # [i f c o] of the input conv, [i f c o] of the recurrent conv
        cst = self._new_buffer(F, Ho, Wo)         # cell state, updated in place
        hseq = self._new_buffer(T * F, Ho, Wo)
        w_in, w_rec = self._weight_id(layer._parts[0]), self._weight_id(layer._parts[1])
        for ts in range(T):
            self._op(nat.OP_CONV, _Val(v.buf, v.c0 + ts * C, C, v.H, v.W), z, dst_c0=0, weight_id=w_in, pad_t=t, pad_b=b,
                     pad_l=l, pad_r=r, pad_mode_h=mh, pad_mode_w=mw, Cout=4 * F, kh=kh, kw=kw, dil_h=dh, dil_w=dw)
            if ts > 0:
                self._op(nat.OP_CONV, _Val(hseq, (ts - 1) * F, F, Ho, Wo), z, dst_c0=4 * F, weight_id=w_rec,
                         pad_t=(kh - 1) // 2, pad_b=kh - 1 - (kh - 1) // 2, pad_l=(kw - 1) // 2,
                         pad_r=kw - 1 - (kw - 1) // 2, pad_mode_h=nat.PAD_ZERO, pad_mode_w=nat.PAD_ZERO, Cout=4 * F,
                         kh=kh, kw=kw)
            self._op(nat.OP_LSTM, _Val(z, 0, 8 * F if ts > 0 else 4 * F, Ho, Wo), hseq, dst_c0=ts * F, Cout=F, act=act,
                     act2=ract, aux=cst, aux_c0=0)
        if layer.return_sequences:
            return _Val(hseq, 0, T * F, Ho, Wo, (), (T, F, Ho, Wo))
        return _Val(hseq, (T - 1) * F, F, Ho, Wo, (), (F, Ho, Wo))

    # -- driver --------------------------------------------------------------------------------------------------
    def _lower(self):
        m = self.model
        in_shape = tuple(m.inputs[0].shape[1:])
        if len(in_shape) not in (3, 4) or any(s is None for s in in_shape):
            raise NotImplementedError('the GPU plan needs a fully defined (C, H, W) or (T, C, H, W) input, got %s' %
                                      (in_shape,))
        fmts = [l.data_format for l in m.layers if hasattr(l, 'data_format')]
        self.fmt = fmts[0] if fmts else 'channels_first'
        if self.fmt == 'channels_last':
            if len(in_shape) != 3:
                raise NotImplementedError('channels_last models are lowered for (H, W, C) inputs only')
            H, W, C = in_shape
            in_buf = self._new_buffer(C, H, W, nat.BUF_INPUT)             # caller memory: (N, H, W, C)
            nchw = self._new_buffer(C, H, W)
            self._op(nat.OP_TO_NCHW, _Val(in_buf, 0, C, H, W), nchw)
            vals = {id(m.inputs[0]): _Val(nchw, 0, C, H, W)}
        else:
            # a (time, channels, lat, lon) input (recurrent nets) is the same memory as (time * channels, lat, lon)
            C, H, W = int(np.prod(in_shape[:-2])), in_shape[-2], in_shape[-1]
            in_buf = self._new_buffer(C, H, W, nat.BUF_INPUT)
            vals = {id(m.inputs[0]): _Val(in_buf, 0, C, H, W, (), in_shape)}
        for node in m._nodes:
            if isinstance(node.layer, InputLayer):
                continue
            vals[id(node.output)] = self._emit(node.layer, [vals[id(t)] for t in node.inputs])
        # model outputs must be dense caller-bound buffers
        used = set()
        for k, t in enumerate(m.outputs):
            v = self._materialise(vals[id(t)])
            b = self.buffers[v.buf]
            full = v.c0 == 0 and v.C == b['C']
            if self.fmt == 'channels_last':
                nb = self._new_buffer(v.C, v.H, v.W, nat.BUF_OUTPUT, k)   # caller memory: (N, H, W, C)
                if not full:
                    tmp = self._new_buffer(v.C, v.H, v.W)
                    self._op(nat.OP_COPY, v, tmp)
                    v = _Val(tmp, 0, v.C, v.H, v.W)
                self._op(nat.OP_TO_NHWC, v, nb)
                v = _Val(nb, 0, v.C, v.H, v.W, (), (v.H, v.W, v.C))
            elif full and b['kind'] == nat.BUF_INTERNAL and v.buf not in used:
                b['kind'], b['output_index'] = nat.BUF_OUTPUT, k
            else:
                nb = self._new_buffer(v.C, v.H, v.W, nat.BUF_OUTPUT, k)
                self._op(nat.OP_COPY, v, nb)
                v = _Val(nb, 0, v.C, v.H, v.W, (), v.shape)
            used.add(v.buf)
            self.out_vals.append(v)
        self._elide_concat_copies()

    def _elide_concat_copies(self):
        """
        A COPY of a whole producer-written INTERNAL buffer into a channel window of another buffer is removed by making
        the producer (and every other reader) use that window directly: concatenate() costs nothing for such inputs.
        """
        changed = True
        while changed:
            changed = False
            for ci, cp in enumerate(self.ops):
                if cp['kind'] != nat.OP_COPY:
                    continue
                sb = self.buffers[cp['src']]
                if sb['kind'] != nat.BUF_INTERNAL or cp['src_c0'] != 0 or cp['src_c'] != sb['C']:
                    continue
                writers = [o for o in self.ops if o['dst'] == cp['src']]
                if len(writers) != 1 or writers[0]['dst_c0'] != 0:
                    continue
                wi = self.ops.index(writers[0])
                # the window must not be written by anything between the producer and the copy
                clash = any(o['dst'] == cp['dst'] for o in self.ops[wi + 1:ci])
                if clash:
                    continue
                src, dst, off = cp['src'], cp['dst'], cp['dst_c0']
                writers[0]['dst'], writers[0]['dst_c0'] = dst, off
                for o in self.ops:
                    if o is not cp and o['src'] == src:
                        o['src'], o['src_c0'] = dst, o['src_c0'] + off
                del self.ops[ci]
                changed = True
                break
        # drop buffers nobody references any more (keep indices stable by compacting)
        live = sorted({o['src'] for o in self.ops} | {o['dst'] for o in self.ops} |
                      {o['aux'] for o in self.ops if o['aux'] >= 0} |
                      {i for i, b in enumerate(self.buffers) if b['kind'] != nat.BUF_INTERNAL})
        remap = {old: new for new, old in enumerate(live)}
        self.buffers = [self.buffers[i] for i in live]
        for o in self.ops:
            o['src'], o['dst'] = remap[o['src']], remap[o['dst']]
            if o['aux'] >= 0:
                o['aux'] = remap[o['aux']]
        for v in self.out_vals:
            v.buf = remap[v.buf]

    def internal_floats_per_sample(self):
        return sum(b['C'] * b['H'] * b['W'] for b in self.buffers if b['kind'] == nat.BUF_INTERNAL)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('dlwp_b200 needs a CUDA device (B200, sm_100a); there is no CPU execution path')
    return torch


class CompiledNet(object):
    """A DlwpPlan plus the host-side glue around it.  Created lazily by keras.Model.engine()."""

    # dlwp_debug_flags bits that void a tensor-core result: NaN / inf data, or a statically scaled tanh image whose amax
    # fell below 2^-6 in scaled units (precision underflow)
    FLAG_TC_FALLBACK = nat.FLAG_TC_RANGE | nat.FLAG_TC_UNDERFLOW

    def __init__(self, model, batch, impl=None, row_windows=None, force_ffma=False, options=None):
        """force_ffma: build the plan on the fp32 kernels (training needs fp32 intermediates).
        row_windows: optional list (one (lo, hi) or None per lowered op) restricting each op to a latitude band
        (dlwp_b200.parallel.BandPlanner.windows); None entries drop the op.
        options: dict of DlwpPlanOptions fields (math, fuse, tc_generic, tc_bands, tc_no_tma, tc_taps_in_k, tc_debug,
        precision: 0 / 'fp32' = fp32-equivalent tensor-core arithmetic, 1 / 'bf16' = plain bf16 images and weights).
        The environment variable DLWP_MATH=ffma is honoured HERE (scripts / bench.py --math), never inside the library."""
        self.torch = _torch()
        self.lib = nat.lib()
        self.model = model
        self.low = Lowering(model)
        self.row_windows = row_windows
        self.options = dict(options or {})
        import os
        self._force_ffma = bool(force_ffma) or (os.environ.get('DLWP_MATH') == 'ffma' and 'math' not in self.options)
        if 'precision' not in self.options and os.environ.get('DLWP_PRECISION', '').lower() == 'bf16':
            self.options['precision'] = nat.PRECISION_BF16      # honoured HERE (scripts / bench.py), never in the library
        if 'latband_spare_sms' not in self.options and os.environ.get('DLWP_LATBAND_SPARE_SMS'):
            self.options['latband_spare_sms'] = int(os.environ['DLWP_LATBAND_SPARE_SMS'])   # -1: no exchange overlap
        if isinstance(self.options.get('precision'), str):
            self.options['precision'] = {'fp32': 0, 'f32': 0, 'bf16': nat.PRECISION_BF16}[self.options['precision'].lower()]
        self.impl = nat.IMPLS[impl] if isinstance(impl, str) else (impl or nat.IMPL_AUTO)
        self.plan = ctypes.c_void_p()
        self.max_batch = 0
        self._pushed = {}
        self.in_shape = tuple(model.inputs[0].shape[1:])   # logical: (C, H, W), or (T, C, H, W) for recurrent nets
        ib = [b for b in self.low.buffers if b['kind'] == nat.BUF_INPUT][0]
        self.in_phys = (ib['C'], ib['H'], ib['W'])
        self.out_shapes = [tuple(v.shape) for v in self.low.out_vals]
        self.out_phys = [(v.C, v.H, v.W) for v in self.low.out_vals]
        self.n_outputs = len(self.out_shapes)
        self._create(self._chunk_for(batch))

    # -- plan management -----------------------------------------------------------------------------------------
    def _chunk_for(self, batch):
        free, _ = self.torch.cuda.mem_get_info()
        per_sample = 4 * max(1, self.low.internal_floats_per_sample())
        cap = max(1, int(0.35 * free) // per_sample)
        return int(max(1, min(batch, cap)))

    def fits(self, batch):
        return batch <= self.max_batch or self._chunk_for(batch) <= self.max_batch

    def _create(self, max_batch):
        bufs = (nat.BufferDesc * len(self.low.buffers))()
        for i, b in enumerate(self.low.buffers):
            bufs[i] = nat.BufferDesc(b['kind'], b['C'], b['H'], b['W'], b['output_index'], 0)
        live = []
        for i, o in enumerate(self.low.ops):
            o = dict(o)
            if o['kind'] == nat.OP_CONV and self.impl:
                o['impl'] = self.impl
            if self.row_windows is not None:
                if self.row_windows[i] is None:
                    continue
                o['row_begin'], o['row_end'] = self.row_windows[i]
            live.append(o)
        ops = (nat.OpDesc * len(live))()
        names = [f[0] for f in nat.OpDesc._fields_]
        for i, o in enumerate(live):
            ops[i] = nat.OpDesc(*[int(o[n]) for n in names])
        net = nat.NetDesc(len(bufs), len(ops), len(self.low.weight_layers), int(max_batch), bufs, ops)
        opts = dict(self.options)
        if self._force_ffma:
            opts['math'] = nat.MATH_FFMA
        po = nat.PlanOptions(**opts)
        nat.check(self.lib.dlwp_plan_create_opts(ctypes.byref(net), ctypes.byref(po), ctypes.byref(self.plan)),
                  'dlwp_plan_create_opts')
        self.max_batch = int(max_batch)
        self._pushed = {}
        self.sync_weights()

    def close(self):
        if self.plan:
            self.lib.dlwp_plan_destroy(self.plan)
            self.plan = ctypes.c_void_p()

    def __del__(self):
        if sys is None or sys.is_finalizing():   # the CUDA context may already be gone at interpreter shutdown
            return
        try:
            self.close()
        except Exception:
            pass

    def sync_weights(self):
        """Push host-side layer weights (Keras layouts) that changed since the last push."""
        for wid, layer in enumerate(self.low.weight_layers):
            if self._pushed.get(wid) == layer._weights_version:
                continue
            k = np.ascontiguousarray(layer._weights[0], dtype=np.float32)
            b = np.ascontiguousarray(layer._weights[1], dtype=np.float32).reshape(-1) if layer.use_bias else None
            nat.check(self.lib.dlwp_plan_set_weights(
                self.plan, wid, k.ctypes.data, k.size, b.ctypes.data if b is not None else None,
                b.size if b is not None else 0), 'dlwp_plan_set_weights')
            self._pushed[wid] = layer._weights_version

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    # -- one model application -----------------------------------------------------------------------------------
    def forward_device(self, x):
        """x: CUDA float32 tensor (N, C, H, W), N <= max_batch.  Returns the list of output tensors (on device)."""
        torch = self.torch
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        n = x.shape[0]
        self.sync_weights()
        outs = [torch.empty((n,) + p, dtype=torch.float32, device=x.device) for p in self.out_phys]
        ptrs = (ctypes.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        nat.check(self.lib.dlwp_plan_forward(self.plan, n, x.data_ptr(), ptrs, self._stream()), 'dlwp_plan_forward')
        return outs

    def forward_into(self, x, outs):
        """Like forward_device but into caller-provided dense output tensors (used by the latitude-band driver)."""
        self.sync_weights()
        ptrs = (ctypes.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        nat.check(self.lib.dlwp_plan_forward(self.plan, x.shape[0], x.data_ptr(), ptrs, self._stream()),
                  'dlwp_plan_forward')

    # -- training -------------------------------------------------------------------------------------------------
    def train_step(self, x, targets, loss_weights=None, backward=True, input_grad=False):
        """Forward + MSE losses (+ backward into the flat gradient buffer).  x, targets: CUDA float32 tensors.
        Returns (per-output mse list, per-output mae list)."""
        n = x.shape[0]
        self.sync_weights()
        k = self.n_outputs
        tp = (ctypes.c_void_p * k)(*[t.data_ptr() for t in targets])
        lw = (ctypes.c_float * k)(*([1.0] * k if loss_weights is None else [float(v) for v in loss_weights]))
        losses, maes = (ctypes.c_float * k)(), (ctypes.c_float * k)()
        nat.check(self.lib.dlwp_train_step(self.plan, n, x.data_ptr(), tp, lw, 1 if backward else 0,
                                           1 if input_grad else 0, losses, maes, self._stream()), 'dlwp_train_step')
        return list(losses), list(maes)

    def _device_view(self, ptr, elems):
        class _Arr(object):
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {'shape': (int(elems),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}
        return self.torch.as_tensor(a, device='cuda')

    def grad_tensor(self):
        """The flat gradient buffer as a CUDA tensor view (for torch.distributed.all_reduce)."""
        p, n, gi = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_void_p()
        nat.check(self.lib.dlwp_train_buffers(self.plan, ctypes.byref(p), ctypes.byref(n), ctypes.byref(gi)))
        return self._device_view(p.value, n.value)

    def input_grad_tensor(self, n):
        p, m, gi = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_void_p()
        nat.check(self.lib.dlwp_train_buffers(self.plan, ctypes.byref(p), ctypes.byref(m), ctypes.byref(gi)))
        return self._device_view(gi.value, n * int(np.prod(self.in_shape))).view((n,) + self.in_shape)

    def weight_grads(self):
        """Gradients in Keras layouts, ordered like model.get_weights() (host numpy)."""
        flat = self.grad_tensor().cpu().numpy()
        out = []
        for wid, layer in enumerate(self.low.weight_layers):
            ko, bo = ctypes.c_int64(), ctypes.c_int64()
            nat.check(self.lib.dlwp_train_weight_offsets(self.plan, wid, ctypes.byref(ko), ctypes.byref(bo)))
            k = layer._weights[0]
            out.append(flat[ko.value:ko.value + k.size].reshape(k.shape).copy())
            if layer.use_bias:
                b = layer._weights[1]
                out.append(flat[bo.value:bo.value + b.size].reshape(b.shape).copy())
        return out

    def adam(self, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        nat.check(self.lib.dlwp_train_adam(self.plan, lr, beta_1, beta_2, epsilon, self._stream()), 'dlwp_train_adam')

    def _regularizers(self):
        """(weight id, kernel l1, kernel l2, bias l1, bias l2) of every layer that carries a keras regularizer."""
        out = []
        for wid, layer in enumerate(self.low.weight_layers):
            kr, br = getattr(layer, 'kernel_regularizer', None), getattr(layer, 'bias_regularizer', None)
            vals = [float(getattr(kr, 'l1', 0.) or 0.), float(getattr(kr, 'l2', 0.) or 0.),
                    float(getattr(br, 'l1', 0.) or 0.), float(getattr(br, 'l2', 0.) or 0.)]
            for r in (kr, br):
                if r is not None and not (hasattr(r, 'l1') and hasattr(r, 'l2')):
                    raise NotImplementedError('only keras.regularizers.l1 / l2 / L1L2 are implemented, got %r' % (r,))
            if any(vals):
                out.append((wid,) + tuple(vals))
        return out

    def regularize(self):
        """Add every layer's L1/L2 term to the flat gradient buffer (call between train_step and adam); returns the
        penalty Keras adds to the loss."""
        total = 0.0
        for wid, kl1, kl2, bl1, bl2 in self._regularizers():
            pen = ctypes.c_float(0)
            nat.check(self.lib.dlwp_train_regularize(self.plan, wid, kl1, kl2, bl1, bl2, ctypes.byref(pen),
                                                     self._stream()), 'dlwp_train_regularize')
            total += float(pen.value)
        return total

    def regularization_penalty(self):
        """The same penalty from the host copy of the weights (evaluation: no gradient to touch)."""
        total = 0.0
        for wid, kl1, kl2, bl1, bl2 in self._regularizers():
            layer = self.low.weight_layers[wid]
            k = np.asarray(layer._weights[0], np.float64)
            total += kl1 * np.abs(k).sum() + kl2 * np.square(k).sum()
            if layer.use_bias:
                b = np.asarray(layer._weights[1], np.float64)
                total += bl1 * np.abs(b).sum() + bl2 * np.square(b).sum()
        return float(total)

    def adam_state(self):
        """(m, v, step) of the optimizer, host copies; None before the first training step."""
        g = ctypes.c_void_p()
        n = ctypes.c_int64()
        if self.lib.dlwp_train_buffers(self.plan, ctypes.byref(g), ctypes.byref(n), None) != 0:
            return None
        m, v, t = np.empty(n.value, np.float32), np.empty(n.value, np.float32), ctypes.c_int64(0)
        nat.check(self.lib.dlwp_train_adam_state(self.plan, m.ctypes.data, v.ctypes.data, n.value, ctypes.byref(t), 0),
                  'dlwp_train_adam_state')
        return m, v, int(t.value)

    def set_adam_state(self, m, v, step):
        t = ctypes.c_int64(int(step))
        nat.check(self.lib.dlwp_train_adam_state(self.plan, m.ctypes.data, v.ctypes.data, m.size, ctypes.byref(t), 1),
                  'dlwp_train_adam_state')

    def pull_weights(self):
        """Device weights -> the front-end layers (after optimizer steps)."""
        for wid, layer in enumerate(self.low.weight_layers):
            k = np.empty(layer._weights[0].shape, np.float32)
            b = np.empty(layer._weights[1].size, np.float32) if layer.use_bias else None
            nat.check(self.lib.dlwp_plan_get_weights(self.plan, wid, k.ctypes.data, k.size,
                                                     b.ctypes.data if b is not None else None,
                                                     b.size if b is not None else 0), 'dlwp_plan_get_weights')
            layer._weights = [k] + ([b.reshape(layer._weights[1].shape)] if b is not None else [])
            layer._weights_version += 1
            self._pushed[wid] = layer._weights_version

    def uses_tensor_cores(self):
        return bool(self.lib.dlwp_plan_uses_tensor_cores(self.plan))

    def fused_pair(self):
        """Index of the first op of the conv -> conv pair that runs as one kernel (csrc/conv_fused.cu), or -1."""
        return int(self.lib.dlwp_plan_fused_pair(self.plan))

    def profile_op(self, n, op_index, iters=20):
        """Average device time (ms) of one op of the plan launched alone (CUDA events); run a forward/rollout first."""
        ms = ctypes.c_float(0)
        nat.check(self.lib.dlwp_plan_profile_op(self.plan, int(n), int(op_index), int(iters), ctypes.byref(ms),
                                                self._stream()), 'dlwp_plan_profile_op')
        return float(ms.value)

    def check_flags(self):
        """Read-and-clear the device flags (synchronises).  True = the last tensor-core results must not be trusted: the
        data held NaN / inf, or a tanh image underflowed its static scale.  The device-resident entry points
        (rollout_device, forward_device, forward_into) are asynchronous and do NOT check; call this after them."""
        return bool(int(self.lib.dlwp_debug_flags()) & self.FLAG_TC_FALLBACK) and self.uses_tensor_cores()

    def _range_fallback(self):
        """
        The tensor-core path stores activations as power-of-two scaled fp16 hi/lo pairs whose exponents follow the data
        (csrc/conv_tc.h), so any finite fp32 magnitude is fine.  What it cannot represent is NaN / inf, and a tanh layer
        whose outputs are ALL below ~1e-6 underflows its static scale: the kernels raise a device flag, the host-level
        entry points then rebuild the plan on the fp32 FFMA kernels and redo the call (True = caller must rerun).
        """
        if not self.check_flags():
            return False
        import warnings
        warnings.warn('dlwp_b200: non-finite activations or a precision underflow on the tensor-core path (fp16-split '
                      'range); falling back to the fp32 FFMA kernels for this model')
        self._force_ffma = True
        mb = self.max_batch
        self.close()
        self._create(mb)
        return True

    def predict(self, x):
        """numpy (N, C, H, W) -> list of numpy outputs (logical shapes), processed in chunks of max_batch."""
        self.lib.dlwp_debug_flags()          # clear stale device flags
        out = self._predict(x)
        if self._range_fallback():
            out = self._predict(x)
        return out

    def _predict(self, x):
        torch = self.torch
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        results = [np.empty((n,) + s, np.float32) for s in self.out_shapes]
        for s in range(0, n, self.max_batch):
            xd = torch.from_numpy(x[s:s + self.max_batch]).cuda()
            outs = self.forward_device(xd)
            for r, o, shp in zip(results, outs, self.out_shapes):
                r[s:s + self.max_batch] = o.cpu().numpy().reshape((-1,) + shp)
        return results

    # -- the rollout ---------------------------------------------------------------------------------------------
    def can_rollout(self):
        return all(p == self.in_phys for p in self.out_phys)

    def rollout_device(self, x0, iterations, use_graph=True, out=None):
        """Device-resident rollout.  x0: CUDA (N,C,H,W).  Returns a CUDA tensor (iterations*n_outputs, N, C, H, W).
        Asynchronous: no range check happens here (NaN / inf inputs give garbage on the tensor-core path); call
        check_flags() afterwards, or use rollout_host / predict, which check and fall back."""
        torch = self.torch
        assert x0.is_cuda and x0.dtype == torch.float32 and x0.is_contiguous()
        n = x0.shape[0]
        if n > self.max_batch:
            raise ValueError('batch %d exceeds the plan capacity %d' % (n, self.max_batch))
        self.sync_weights()
        shape = (iterations * self.n_outputs, n) + self.in_shape
        series = out if out is not None else torch.empty(shape, dtype=torch.float32, device=x0.device)
        assert tuple(series.shape) == shape and series.is_contiguous()
        nat.check(self.lib.dlwp_rollout(self.plan, n, x0.data_ptr(), series.data_ptr(), int(iterations),
                                        1 if use_graph else 0, self._stream()), 'dlwp_rollout')
        return series

    def rollout_step_sequence_host(self, x0, iterations, time_dim):
        """predict_timeseries(step_sequence=True) as one device-resident loop (dlwp_rollout_step_sequence): numpy in,
        numpy (iterations, N) + input shape out; batches beyond the plan capacity run as sample chunks."""
        torch = self.torch
        x0 = np.ascontiguousarray(x0, dtype=np.float32)
        n = x0.shape[0]
        self.sync_weights()
        out = np.empty((iterations, n) + self.in_shape, np.float32)
        for s in range(0, n, self.max_batch):
            for attempt in (0, 1):
                self.lib.dlwp_debug_flags()
                xd = torch.from_numpy(x0[s:s + self.max_batch]).cuda()
                series = torch.empty((iterations, xd.shape[0]) + self.in_shape, dtype=torch.float32, device='cuda')
                nat.check(self.lib.dlwp_rollout_step_sequence(self.plan, xd.shape[0], xd.data_ptr(), series.data_ptr(),
                                                              int(iterations), int(time_dim), self._stream()),
                          'dlwp_rollout_step_sequence')
                out[:, s:s + self.max_batch] = series.cpu().numpy()
                if not self._range_fallback():
                    break
        return out

    def rollout_host(self, x0, iterations, d2h_group=0, pinned=True):
        self.lib.dlwp_debug_flags()          # clear stale device flags
        out = self._rollout_host(x0, iterations, d2h_group, pinned)
        if self._range_fallback():
            out = self._rollout_host(x0, iterations, d2h_group, pinned)
        return out

    def _rollout_host(self, x0, iterations, d2h_group=0, pinned=True):
        """
        numpy in -> numpy out through dlwp_rollout_host: H2D of x0, rollout, D2H of the series pipelined behind the
        compute.  The result lives in pinned host memory from torch's caching host allocator (a fresh array per call).
        Batches larger than the plan capacity run as independent sample chunks.
        """
        torch = self.torch
        x0 = np.ascontiguousarray(x0, dtype=np.float32)
        n = x0.shape[0]
        self.sync_weights()
        slots = iterations * self.n_outputs
        shape = (slots, n) + self.in_shape
        if n <= self.max_batch:
            buf = torch.empty(shape, dtype=torch.float32, pin_memory=pinned)
            series = buf.numpy()
            nat.check(self.lib.dlwp_rollout_host(self.plan, n, x0.ctypes.data, series.ctypes.data, int(iterations),
                                                 int(d2h_group)), 'dlwp_rollout_host')
            return series
        series = np.empty(shape, np.float32)
        for s in range(0, n, self.max_batch):
            part = self._rollout_host(x0[s:s + self.max_batch], iterations, d2h_group, pinned)
            series[:, s:s + self.max_batch] = part
        return series
