"""
GPU parity of the single-layer kernels, called through the C ABI, against the float64 oracle.

Tolerance: the kernels accumulate in fp32 (FFMA) in a different order than the oracle's float64 sum, so the bar is
max|gpu - oracle| <= 2e-5 * max|oracle| per layer (observed ~1e-6); data-movement ops are bit exact.
"""

import ctypes
import itertools

import numpy as np
import pytest

from oracle import ops as OO
from tests.helpers import rel_err, run_conv

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.fixture(scope='module')
def env():
    import torch
    from dlwp_b200 import _native as nat
    nat.lib()
    return nat, torch


def _oracle(x, k, b, d, pads, mh, mw, act):
    y = OO.pad_conv2d_closed_form(x.astype(np.float64), k.astype(np.float64), None if b is None else b.astype(np.float64),
                                  (d, d), pads[0], pads[1], 'periodic' if mh else 'zero', 'periodic' if mw else 'zero')
    return OO.activation({0: None, 1: 'tanh', 2: 'relu'}[act])(y)


EXAMPLE_LAYERS = [
    # (Cin, Cout, k, d, pad)   -- every conv of the example nets, on a reduced grid
    (6, 32, 3, 2, 2), (32, 6, 5, 1, 2), (12, 32, 3, 2, 2), (16, 64, 3, 1, 1), (32, 128, 3, 1, 1), (128, 32, 3, 1, 1),
    (64, 16, 3, 2, 2), (32, 12, 5, 1, 2), (128, 64, 3, 1, 1), (64, 32, 3, 2, 2),
]


@pytest.mark.parametrize('impl', ['direct', 'ffma', 'ffma_tma', 'auto'])
@pytest.mark.parametrize('layer', EXAMPLE_LAYERS)
def test_example_layers_periodic_lon_zero_lat(env, impl, layer):
    nat, torch = env
    cin, cout, k, d, pad = layer
    rng = np.random.RandomState(cin * 1000 + cout)
    x = rng.standard_normal((3, cin, 23, 36)).astype(np.float32)
    w = OO.glorot_uniform(rng, k, k, cin, cout)
    b = (0.1 * rng.standard_normal(cout)).astype(np.float32)
    act = nat.ACT_LINEAR if cout in (6, 12) else nat.ACT_TANH
    pads = ((pad, pad), (pad, pad))
    y = run_conv(nat, torch, x, w, b, d, pads, nat.PAD_ZERO, nat.PAD_PERIODIC, act, nat.IMPLS[impl])
    ref = _oracle(x, w, b, d, pads, 0, 1, act)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < TOL


@pytest.mark.parametrize('impl', ['direct', 'ffma', 'ffma_tma'])
def test_full_size_net_a_layers(env, impl):
    """BASELINE.json configs[1] shapes: 6x91x180 -> 32 (k3 d2 tanh) -> 6 (k5)."""
    nat, torch = env
    rng = np.random.RandomState(5)
    x = rng.standard_normal((2, 6, 91, 180)).astype(np.float32)
    w1, b1 = OO.glorot_uniform(rng, 3, 3, 6, 32), (0.1 * rng.standard_normal(32)).astype(np.float32)
    w2, b2 = OO.glorot_uniform(rng, 5, 5, 32, 6), (0.1 * rng.standard_normal(6)).astype(np.float32)
    pads = ((2, 2), (2, 2))
    h = run_conv(nat, torch, x, w1, b1, 2, pads, nat.PAD_ZERO, nat.PAD_PERIODIC, nat.ACT_TANH, nat.IMPLS[impl])
    href = _oracle(x, w1, b1, 2, pads, 0, 1, nat.ACT_TANH)
    assert rel_err(h, href) < TOL
    y = run_conv(nat, torch, h, w2, b2, 1, pads, nat.PAD_ZERO, nat.PAD_PERIODIC, nat.ACT_LINEAR, nat.IMPLS[impl])
    assert rel_err(y, _oracle(h, w2, b2, 1, pads, 0, 1, nat.ACT_LINEAR)) < TOL


@pytest.mark.parametrize('impl', ['direct', 'ffma', 'ffma_tma'])
def test_wide_grid_uses_two_tiles_along_longitude(env, impl):
    """W = 360 (1-degree grid): rows no longer fit one TMA box, both edge tiles need their own wrap patch."""
    nat, torch = env
    rng = np.random.RandomState(6)
    x = rng.standard_normal((1, 12, 20, 360)).astype(np.float32)
    w, b = OO.glorot_uniform(rng, 5, 5, 12, 12), (0.1 * rng.standard_normal(12)).astype(np.float32)
    pads = ((2, 2), (2, 2))
    y = run_conv(nat, torch, x, w, b, 1, pads, nat.PAD_ZERO, nat.PAD_PERIODIC, nat.ACT_LINEAR, nat.IMPLS[impl])
    assert rel_err(y, _oracle(x, w, b, 1, pads, 0, 1, nat.ACT_LINEAR)) < TOL


PAD_CASES = list(itertools.product([(3, 1), (3, 2), (5, 1)], [0, 1], [0, 1]))


@pytest.mark.parametrize('impl', ['direct', 'ffma', 'auto'])
@pytest.mark.parametrize('kd,mh,mw', PAD_CASES)
def test_pad_modes_and_asymmetric_pads(env, impl, kd, mh, mw):
    """Every combination of zero / periodic per axis, asymmetric ((t,b),(l,r)), odd sizes, ragged tile edges."""
    nat, torch = env
    k, d = kd
    rng = np.random.RandomState(k * 10 + d + 2 * mh + mw)
    for (N, cin, H, W, cout, pads) in [(2, 5, 13, 18, 7, ((1, 3), (2, 0))), (1, 3, 9, 30, 4, ((4, 0), (0, 4))),
                                       (2, 8, 17, 21, 8, ((2, 2), (3, 1)))]:
        if (k - 1) * d > min(H + sum(pads[0]), W + sum(pads[1])) - 1:
            continue
        x = rng.standard_normal((N, cin, H, W)).astype(np.float32)
        w = OO.glorot_uniform(rng, k, k, cin, cout)
        b = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        y = run_conv(nat, torch, x, w, b, d, pads, mh, mw, nat.ACT_RELU, nat.IMPLS[impl])
        assert rel_err(y, _oracle(x, w, b, d, pads, mh, mw, nat.ACT_RELU)) < TOL


def test_no_bias_and_general_kernel_sizes_fall_back_to_direct(env):
    nat, torch = env
    rng = np.random.RandomState(8)
    x = rng.standard_normal((2, 4, 10, 12)).astype(np.float32)
    for (kh, kw, d) in [(1, 1, 1), (3, 5, 1), (7, 7, 1), (3, 3, 3)]:
        w = rng.standard_normal((kh, kw, 4, 5)).astype(np.float32) * 0.2
        th, tw = d * (kh - 1), d * (kw - 1)
        pads = ((th // 2, th - th // 2), (tw // 2, tw - tw // 2))
        y = run_conv(nat, torch, x, w, None, d, pads, nat.PAD_ZERO, nat.PAD_PERIODIC, nat.ACT_LINEAR, nat.IMPL_AUTO)
        assert rel_err(y, _oracle(x, w, None, d, pads, 0, 1, nat.ACT_LINEAR)) < TOL


def test_pool_and_upsample_on_load(env):
    nat, torch = env
    rng = np.random.RandomState(9)
    x = rng.standard_normal((2, 6, 12, 16)).astype(np.float32)
    w, b = OO.glorot_uniform(rng, 3, 3, 6, 8), (0.1 * rng.standard_normal(8)).astype(np.float32)
    pads = ((1, 1), (1, 1))
    for pre, f in ((1, OO.max_pool2d), (2, OO.upsample2d)):
        y = run_conv(nat, torch, x, w, b, 1, pads, nat.PAD_ZERO, nat.PAD_PERIODIC, nat.ACT_TANH, nat.IMPL_AUTO, pre_op=pre)
        ref = _oracle(f(x.astype(np.float64), 2), w, b, 1, pads, 0, 1, nat.ACT_TANH)
        assert rel_err(y, ref) < TOL


def test_row_connected_matches_oracle_and_reference_golden(env, golden_dir):
    import os
    nat, torch = env
    g = np.load(os.path.join(golden_dir, 'row_conv2d.npz'))
    x, k = g['x'].astype(np.float32), g['kernel'].astype(np.float32)
    y = run_conv(nat, torch, x, k, None, 1, ((0, 0), (0, 0)), 0, 0, nat.ACT_LINEAR, nat.IMPL_AUTO, rowwise=1)
    assert rel_err(y, g['y']) < TOL                        # output of the reference's own row_conv2d
    rng = np.random.RandomState(10)
    bias = (0.1 * rng.standard_normal((5, 1, 3))).astype(np.float32)
    yb = run_conv(nat, torch, x, k, bias.reshape(5, 3), 1, ((0, 0), (0, 0)), 0, 0, nat.ACT_TANH, nat.IMPL_AUTO, rowwise=1)
    ref = np.tanh(OO.row_conv2d(g['x'], g['kernel'], bias.astype(np.float64)))
    assert rel_err(yb, ref) < TOL


@pytest.mark.parametrize('k,d,cout', [(5, 1, 12), (3, 1, 6), (3, 2, 16), (5, 1, 3)])
def test_row_connected_tiled_kernel_equals_direct_kernel(env, k, d, cout):
    """conv_rowwise_kernel (one CTA per output row, the row's weights in shared memory) against conv_direct_kernel and the
    oracle: periodic longitude + zero latitude padding as in examples/train_functional.py:191-196, ragged width."""
    nat, torch = env
    rng = np.random.RandomState(12)
    N, Cin, H, W = 3, 9, 11, 30
    pad = d * (k - 1) // 2
    x = rng.standard_normal((N, Cin, H, W)).astype(np.float32)
    kern = (0.2 * rng.standard_normal((H, k, k, Cin, cout))).astype(np.float32)
    bias = (0.1 * rng.standard_normal((H, cout))).astype(np.float32)
    pads = ((pad, pad), (pad, pad))
    args = (nat, torch, x, kern, bias, d, pads, nat.PAD_ZERO, nat.PAD_PERIODIC, nat.ACT_TANH)
    y_tiled = run_conv(*args, nat.IMPL_AUTO, rowwise=1)
    y_direct = run_conv(*args, nat.IMPL_DIRECT, rowwise=1)
    assert y_tiled.shape == (N, cout, H, W)
    assert np.abs(y_tiled - y_direct).max() < 1e-5       # fp32, different summation order
    xp = OO.zero_pad2d(OO.periodic_pad2d(x.astype(np.float64), ((0, 0), (pad, pad))), ((pad, pad), (0, 0)))
    ref = np.stack([OO.conv2d_valid(xp[:, :, r:r + d * (k - 1) + 1], kern[r].astype(np.float64), None, (d, d))[:, :, 0]
                    for r in range(H)], axis=2) + bias.astype(np.float64).T[None, :, :, None]
    assert rel_err(y_tiled, np.tanh(ref)) < TOL


def test_elementwise_ops_bit_exact(env, golden_dir):
    import os
    nat, torch = env
    lib = nat.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = np.load(os.path.join(golden_dir, 'periodic_padding2d.npz'))
    x = g['x_channels_first']
    N, C, H, W = x.shape
    xd = torch.from_numpy(x).cuda()
    for kcase in range(int(g['n_cases'])):
        (t, b), (l, r) = g['pad_%d' % kcase]
        ref = g['y_%d_channels_first' % kcase]                 # the reference's own PeriodicPadding2D.call
        Ho, Wo = H + t + b, W + l + r
        yd = torch.empty((N, C, Ho, Wo), dtype=torch.float32, device='cuda')
        nat.check(lib.dlwp_pad2d(xd.data_ptr(), yd.data_ptr(), N, C, H, W, int(t), int(b), int(l), int(r), 1, 1,
                                 C * H * W, H * W, W, C * Ho * Wo, Ho * Wo, Wo, stream))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(yd.cpu().numpy(), ref)
        nat.check(lib.dlwp_pad2d(xd.data_ptr(), yd.data_ptr(), N, C, H, W, int(t), int(b), int(l), int(r), 0, 0,
                                 C * H * W, H * W, W, C * Ho * Wo, Ho * Wo, Wo, stream))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(yd.cpu().numpy(), OO.zero_pad2d(x, ((t, b), (l, r))))
    x2 = np.random.RandomState(3).standard_normal((2, 3, 9, 10)).astype(np.float32)   # odd H: floor
    xd = torch.from_numpy(x2).cuda()
    yd = torch.empty((2, 3, 4, 5), dtype=torch.float32, device='cuda')
    nat.check(lib.dlwp_maxpool2d(xd.data_ptr(), yd.data_ptr(), 2, 3, 9, 10, 270, 90, 10, 60, 20, 5, stream))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(yd.cpu().numpy(), OO.max_pool2d(x2, 2))
    yd = torch.empty((2, 3, 18, 20), dtype=torch.float32, device='cuda')
    nat.check(lib.dlwp_upsample2d(xd.data_ptr(), yd.data_ptr(), 2, 3, 9, 10, 270, 90, 10, 1080, 360, 20, stream))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(yd.cpu().numpy(), OO.upsample2d(x2, 2))


def test_argument_errors_are_reported_not_thrown(env):
    nat, torch = env
    from tests.helpers import conv_desc
    lib = nat.lib()
    x = torch.zeros((1, 2, 4, 4), device='cuda')
    desc, _, _ = conv_desc(nat, 1, 2, 4, 4, 2, 3, 3, 1, ((1, 1), (1, 1)), 0, 1, 0, 0)
    assert lib.dlwp_conv2d_fwd(ctypes.byref(desc), None, x.data_ptr(), None, x.data_ptr(), None) == -1
    assert b'null' in lib.dlwp_last_error_string()
    desc.pad_l = 9                                           # periodic pad wider than the axis (SURVEY.md A.7 iii)
    assert lib.dlwp_conv2d_fwd(ctypes.byref(desc), x.data_ptr(), x.data_ptr(), None, x.data_ptr(), None) == -2
    desc.pad_l = 1
    desc.act = 7
    assert lib.dlwp_conv2d_fwd(ctypes.byref(desc), x.data_ptr(), x.data_ptr(), None, x.data_ptr(), None) == -1
    assert lib.dlwp_plan_create(None, None) == -1
