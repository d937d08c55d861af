"""Shared builders: the product model (dlwp_b200) and the oracle net (oracle/) with identical weights."""

import ctypes

import numpy as np

from oracle import layers as OL
from oracle import rollout as OR


def build_product_sequential(layer_tuples, time_dim=1, **wrapper_kwargs):
    from dlwp_b200.model import DLWPNeuralNet
    kw = dict(is_convolutional=True, is_recurrent=False, time_dim=time_dim, scaler_type=None, scale_targets=False)
    kw.update(wrapper_kwargs)
    dlwp = DLWPNeuralNet(**kw)
    dlwp.build_model(layer_tuples, loss='mse', optimizer='adam')
    return dlwp


def oracle_sequential_like(dlwp, layer_tuples, seed=1, bias_scale=0.05):
    """Oracle net with seeded weights; the same weights are pushed into the product model."""
    net = OL.OSequential(layer_tuples)
    OL.init_weights(net.conv_layers, seed=seed, bias_scale=bias_scale)
    dlwp.model.set_weights(net.get_weights())
    return net


def build_functional_pair(cs, skip=True, integration_steps=1, seed=1, bias_scale=0.05, latitude_dependent=False):
    """examples/train_functional.py:154-285 built through the product's keras front end + the oracle twin."""
    from dlwp_b200 import keras
    from dlwp_b200.custom import PeriodicPadding2D, RowConnected2D, slice_layer
    from dlwp_b200.keras.layers import Conv2D, Input, MaxPooling2D, UpSampling2D, ZeroPadding2D, concatenate
    from dlwp_b200.model import DLWPFunctional
    cf = 'channels_first'
    x0 = Input(shape=cs, name='input_0')
    pp2, zp2 = PeriodicPadding2D(padding=(0, 2), data_format=cf), ZeroPadding2D(padding=(2, 0), data_format=cf)
    pp1, zp1 = PeriodicPadding2D(padding=(0, 1), data_format=cf), ZeroPadding2D(padding=(1, 0), data_format=cf)
    pool, up = MaxPooling2D(2, data_format=cf), UpSampling2D(2, data_format=cf)
    kw = {'padding': 'valid', 'activation': 'tanh', 'data_format': cf}
    c1 = Conv2D(32, 3, dilation_rate=2, **kw)
    c2 = Conv2D(64, 3, dilation_rate=1, **kw)
    c3 = Conv2D(128, 3, dilation_rate=1, **kw)
    c4 = Conv2D(32 if skip else 64, 3, dilation_rate=1, **kw)
    c5 = Conv2D(16 if skip else 32, 3, dilation_rate=2, **kw)
    last = RowConnected2D if latitude_dependent else Conv2D
    c6 = last(cs[0], 5, padding='valid', activation='linear', data_format=cf)
    s11, s12 = slice_layer(0, 16, axis=1), slice_layer(16, 32, axis=1)
    s21, s22 = slice_layer(0, 32, axis=1), slice_layer(32, 64, axis=1)

    def basic(x):
        x = pool(c1(pp2(zp2(x))))
        x = pool(c2(pp1(zp1(x))))
        x = up(c3(pp1(zp1(x))))
        x = up(c4(pp1(zp1(x))))
        x = c5(pp2(zp2(x)))
        return c6(pp2(zp2(x)))

    def skipm(x):
        x = c1(pp2(zp2(x)))
        x, x1 = s11(x), s12(x)
        x = pool(x)
        x = c2(pp1(zp1(x)))
        x, x2 = s21(x), s22(x)
        x = pool(x)
        x = c3(pp1(zp1(x)))
        x = up(x)
        x = c4(pp1(zp1(x)))
        x = concatenate([x, x2], axis=1)
        x = up(x)
        x = c5(pp2(zp2(x)))
        x = concatenate([x, x1], axis=1)
        return c6(pp2(zp2(x)))

    f = skipm if skip else basic
    outs = [f(x0)]
    for _ in range(1, integration_steps):
        outs.append(f(outs[-1]))
    model = keras.Model(inputs=x0, outputs=outs)
    dlwp = DLWPFunctional(is_convolutional=True, is_recurrent=False, time_dim=1)
    dlwp.build_model(model, loss='mse', optimizer='adam', loss_weights=[1. / integration_steps] * integration_steps)
    onet = OL.OFunctionalNet(cs, skip_connections=skip, integration_steps=integration_steps,
                             latitude_dependent=latitude_dependent)
    OL.init_weights(onet.conv_layers, seed=seed, bias_scale=bias_scale)
    for layer, ol in zip((c1, c2, c3, c4, c5, c6), onet.conv_layers):
        layer.set_weights(ol.weights)
    return dlwp, onet


def rel_err(a, ref):
    """max|a - ref| / max|ref| -- the metric of BASELINE.json's parity gate."""
    return float(np.abs(np.asarray(a, np.float64) - ref).max() / np.abs(ref).max())


def oracle_rollout64(net, x0, steps):
    return OR.neuralnet_predict_timeseries(lambda p: net.forward(p), np.asarray(x0, np.float64), steps,
                                           dtype=np.float64)


def conv_desc(nat, N, Cin, H, W, Cout, kh, kw, d, pads, mode_h, mode_w, act, impl, pre_op=0, rowwise=0):
    """Dense-tensor DlwpConvDesc."""
    (pt, pb), (pl, pr) = pads
    Hl, Wl = (H // 2, W // 2) if pre_op == 1 else ((H * 2, W * 2) if pre_op == 2 else (H, W))
    Ho = Hl + pt + pb - d * (kh - 1)
    Wo = Wl + pl + pr - d * (kw - 1)
    desc = nat.ConvDesc(N=N, Cin=Cin, H=H, W=W, Cout=Cout, kh=kh, kw=kw, dil_h=d, dil_w=d, pad_t=pt, pad_b=pb,
                        pad_l=pl, pad_r=pr, pad_mode_h=mode_h, pad_mode_w=mode_w, act=act, pre_op=pre_op,
                        rowwise=rowwise, impl=impl, reserved=0, row_begin=0, row_end=0, x_stride_n=Cin * H * W, x_stride_c=H * W, x_stride_h=W,
                        y_stride_n=Cout * Ho * Wo, y_stride_c=Ho * Wo, y_stride_h=Wo)
    return desc, Ho, Wo


def run_conv(nat, torch, x, k, b, d, pads, mode_h, mode_w, act, impl, pre_op=0, rowwise=0):
    """Call dlwp_conv2d_fwd through the C ABI on numpy inputs; returns numpy output."""
    N, Cin, H, W = x.shape
    kh, kw = k.shape[-4], k.shape[-3]
    Cout = k.shape[-1]
    desc, Ho, Wo = conv_desc(nat, N, Cin, H, W, Cout, kh, kw, d, pads, mode_h, mode_w, act, impl, pre_op, rowwise)
    xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
    kd = torch.from_numpy(np.ascontiguousarray(k, np.float32)).cuda()
    bd = torch.from_numpy(np.ascontiguousarray(b, np.float32)).cuda() if b is not None else None
    yd = torch.full((N, Cout, Ho, Wo), float('nan'), dtype=torch.float32, device='cuda')
    rc = nat.lib().dlwp_conv2d_fwd(ctypes.byref(desc), xd.data_ptr(), kd.data_ptr(),
                                   bd.data_ptr() if bd is not None else None, yd.data_ptr(),
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    nat.check(rc, 'dlwp_conv2d_fwd')
    torch.cuda.synchronize()
    return yd.cpu().numpy()


def bf16_round(x):
    """Round-to-nearest-even to bfloat16, returned as float64 (what the bf16 mode stores for activations and weights)."""
    a = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    a = (a + 0x7FFF + ((a >> 16) & 1)) & 0xFFFF0000
    return a.astype(np.uint32).view(np.float32).astype(np.float64).reshape(np.shape(x))
