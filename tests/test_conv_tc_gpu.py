"""
Tensor-core (tcgen05) conv path vs the float64 oracle, per layer through the C ABI (impl = DLWP_IMPL_TC).
Operands are fp16 hi/lo splits (3 MMAs per K step, fp32 accumulation in TMEM): the bar stays 2e-5 of max|oracle|.
"""

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


def _check(nat, torch, N, cin, H, W, cout, k, d, act, seed):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((N, cin, H, W)).astype(np.float32)
    w = OO.glorot_uniform(rng, k, k, cin, cout)
    b = (0.1 * rng.standard_normal(cout)).astype(np.float32)
    pad = d * (k - 1) // 2
    pads = ((pad, pad), (pad, pad))
    y = run_conv(nat, torch, x, w, b, d, pads, nat.PAD_ZERO, nat.PAD_PERIODIC, act, nat.IMPL_TC)
    assert nat.lib().dlwp_debug_flags() == 0
    ref = OO.pad_conv2d_closed_form(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (d, d), pads[0],
                                    pads[1], 'zero', 'periodic')
    ref = OO.activation({0: None, 1: 'tanh', 2: 'relu'}[act])(ref)
    err = rel_err(y, ref)
    assert err < TOL, err
    return err


def test_net_a_conv2_32_to_6_k5(env):
    nat, torch = env
    _check(nat, torch, 2, 32, 91, 180, 6, 5, 1, nat.ACT_LINEAR, 1)


def test_net_a_conv1_6_to_32_k3_d2_tanh(env):
    nat, torch = env
    _check(nat, torch, 2, 6, 91, 180, 32, 3, 2, nat.ACT_TANH, 2)


@pytest.mark.parametrize('case', [(3, 6, 23, 36, 32, 3, 2), (3, 32, 23, 36, 6, 5, 1), (2, 16, 17, 44, 64, 3, 1),
                                  (2, 64, 12, 60, 16, 3, 2), (2, 12, 14, 40, 12, 5, 1),
                                  (1, 128, 10, 48, 32, 3, 1), (5, 8, 7, 124, 8, 3, 1)])
def test_other_geometries(env, case):
    nat, torch = env
    N, cin, H, W, cout, k, d = case
    _check(nat, torch, N, cin, H, W, cout, k, d, nat.ACT_RELU, sum(case))


def test_many_tiles_persistent_loop(env):
    """More tiles than SMs: every CTA loops, TMEM accumulator sets and smem stages wrap their phases several times."""
    nat, torch = env
    _check(nat, torch, 24, 32, 91, 180, 6, 5, 1, nat.ACT_LINEAR, 7)


def test_net_a_rollout_on_tensor_cores_meets_the_gate(env):
    """The whole plan as a tensor-core chain: state kept in P layout between layers AND between iterations (the last conv
    re-packs the next input), fp32 series written every step."""
    nat, torch = env
    from dlwp_b200.engine import CompiledNet
    from oracle import layers as OL
    from tests.helpers import build_product_sequential, oracle_rollout64, oracle_sequential_like
    layers = OL.net_a_layers()
    dlwp = build_product_sequential(layers)
    net = oracle_sequential_like(dlwp, layers, seed=1, bias_scale=0.02)
    x0 = np.random.RandomState(0).standard_normal((3, 6, 91, 180)).astype(np.float32)
    eng = CompiledNet(dlwp.model, 3, impl='tc')
    assert eng.uses_tensor_cores()
    xd = torch.from_numpy(x0).cuda()
    got = eng.rollout_device(xd, 50, use_graph=True).cpu().numpy()
    assert nat.lib().dlwp_debug_flags() == 0
    ref = oracle_rollout64(net, x0, 50)
    per_step = [rel_err(got[t], ref[t]) for t in range(50)]
    assert max(per_step) <= 1e-4, per_step
    plain = eng.rollout_device(xd, 50, use_graph=False).cpu().numpy()
    np.testing.assert_array_equal(plain, got)
    one = eng.predict(x0)[0]                                  # single application (packs the input itself)
    np.testing.assert_array_equal(one, got[0])
    host = eng.rollout_host(x0, 7)
    np.testing.assert_array_equal(host, got[:7])
    ffma = dlwp.predict_timeseries(x0, 50)                    # the fp32 FFMA path agrees to ~1e-6
    assert rel_err(ffma, got.astype(np.float64)) < 2e-5
    eng.close()


def test_unrolled_two_step_model_on_tensor_cores(env):
    """n_outputs = 2 (shared weights applied twice): the first output is both a model output and a conv input."""
    nat, torch = env
    from dlwp_b200 import keras
    from dlwp_b200.engine import CompiledNet
    from oracle import layers as OL
    from oracle import rollout as OR
    from tests.helpers import build_product_sequential, oracle_sequential_like
    layers = OL.net_a_layers((6, 30, 64))
    seq = build_product_sequential(layers)
    net = oracle_sequential_like(seq, layers, seed=5, bias_scale=0.05)
    x_in = keras.Input(shape=(6, 30, 64))

    def apply(t):
        for layer in seq.model.layers:
            t = layer(t)
        return t
    o1 = apply(x_in)
    model = keras.Model(inputs=x_in, outputs=[o1, apply(o1)])
    eng = CompiledNet(model, 2, impl='tc')
    assert eng.uses_tensor_cores()
    x0 = np.random.RandomState(3).standard_normal((2, 6, 30, 64)).astype(np.float32)
    got = eng.rollout_host(x0, 3)

    def fn(p):
        a = net.forward(p)
        return [a, net.forward(a)]
    ref = OR.functional_predict_timeseries(fn, x0.astype(np.float64), 6, n_steps=2, dtype=np.float64)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 2e-5
    eng.close()


def test_values_beyond_fp16_range_stay_on_tensor_cores(env):
    """|x| > 65504 used to overflow the unscaled fp16 split (round 1 fell back to FFMA); the image's exponent now follows
    the measured amax, so raw-magnitude fields (geopotential ~ 5e4 .. 3e5) run on the tensor cores at full accuracy."""
    nat, torch = env
    import warnings
    from dlwp_b200.engine import CompiledNet
    from oracle import layers as OL
    from tests.helpers import build_product_sequential, oracle_sequential_like
    layers = OL.net_a_layers((6, 20, 36))
    dlwp = build_product_sequential(layers)
    net = oracle_sequential_like(dlwp, layers, seed=2, bias_scale=0.0)
    x0 = (3.0e5 * np.random.RandomState(1).standard_normal((2, 6, 20, 36))).astype(np.float32)
    eng = CompiledNet(dlwp.model, 2)
    assert eng.uses_tensor_cores()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        y = eng.predict(x0)[0]
    assert not w and eng.uses_tensor_cores()
    assert rel_err(y, net.forward(x0.astype(np.float64))) < 2e-5
    eng.close()


def test_non_finite_inputs_fall_back_to_fp32_kernels(env):
    """NaN / inf cannot be split: the kernels flag it and the engine reruns on the FFMA path (which propagates them like
    the reference's fp32 arithmetic)."""
    nat, torch = env
    import warnings
    from dlwp_b200.engine import CompiledNet
    from oracle import layers as OL
    from tests.helpers import build_product_sequential, oracle_sequential_like
    layers = OL.net_a_layers((6, 20, 36))
    dlwp = build_product_sequential(layers)
    oracle_sequential_like(dlwp, layers, seed=2, bias_scale=0.0)
    x0 = np.random.RandomState(1).standard_normal((2, 6, 20, 36)).astype(np.float32)
    x0[0, 0, 3, 5] = np.nan
    eng = CompiledNet(dlwp.model, 2)
    assert eng.uses_tensor_cores()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        y = eng.predict(x0)[0]
    assert any('fp16-split range' in str(m.message) for m in w)
    assert not eng.uses_tensor_cores()
    assert not np.isfinite(y[0]).all() and np.isfinite(y[1]).all()      # sample 1 is untouched by sample 0's NaN
    eng.close()


def test_latitude_band_windows_on_tensor_cores(env):
    """Row-windowed tensor-core plans (one per band) reproduce the single-domain tensor-core rollout to round-off (the
    band-local amax gives the re-packed state a band-local exponent; see tests/test_latband_gpu.py)."""
    nat, torch = env
    from oracle import layers as OL
    from tests.helpers import build_product_sequential, oracle_sequential_like
    from tests.test_latband_gpu import _run_bands
    layers = OL.net_a_layers()
    dlwp = build_product_sequential(layers)
    oracle_sequential_like(dlwp, layers, seed=1, bias_scale=0.05)
    x0 = np.random.RandomState(0).standard_normal((2, 6, 91, 180)).astype(np.float32)
    assert dlwp.model.engine(2).uses_tensor_cores()
    ref = dlwp.predict_timeseries(x0, 4)
    got, _ = _run_bands(dlwp.model, 4, x0, 4)
    assert nat.lib().dlwp_debug_flags() == 0                  # the NaN-poisoned rows outside band + halo were never read
    assert np.isfinite(got).all() and np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max()


def test_one_degree_grid_rows_wider_than_a_tma_box(env):
    """W = 360 (the 1-degree grid): rows are staged whole by bulk copies, no 256-element box limit."""
    nat, torch = env
    _check(nat, torch, 1, 12, 22, 360, 32, 3, 2, nat.ACT_TANH, 11)
    _check(nat, torch, 1, 32, 16, 360, 12, 5, 1, nat.ACT_LINEAR, 12)


@pytest.mark.parametrize('cs,N', [((12, 24, 48), 3), ((12, 36, 104), 2)])
def test_unet_skip_model_runs_as_a_tensor_core_chain(env, cs, N):
    """Net B (examples/train_functional.py:248-275): convs on tcgen05, MaxPooling2D / UpSampling2D / skip-connection
    copies as data movers on P images, slice_layer / concatenate as plane windows.  fp16 hi/lo split keeps fp32-level
    accuracy: same 5e-5 bar as the fp32 path; a 6-step rollout (feedback re-packed by the last conv) stays under 1e-4."""
    nat, torch = env
    from dlwp_b200.engine import CompiledNet
    from oracle import rollout as OR
    from tests.helpers import build_functional_pair
    dlwp, onet = build_functional_pair(cs, skip=True, integration_steps=1, seed=3)
    x0 = np.random.RandomState(4).standard_normal((N,) + cs).astype(np.float32)
    eng = CompiledNet(dlwp.model, N, impl='tc')
    assert eng.uses_tensor_cores()
    y = eng.predict(x0)[0]
    assert nat.lib().dlwp_debug_flags() == 0
    ref = onet.forward(x0.astype(np.float64))
    assert rel_err(y, ref) < 5e-5
    got = eng.rollout_host(x0, 6)
    refs = OR.functional_predict_timeseries(lambda p: onet.forward(p), x0.astype(np.float64), 6, n_steps=1, dtype=np.float64)
    assert got.shape == refs.shape
    assert rel_err(got, refs) < 1e-4
    xd = torch.from_numpy(x0).cuda()
    g1 = eng.rollout_device(xd, 4, use_graph=True).cpu().numpy()
    g0 = eng.rollout_device(xd, 4, use_graph=False).cpu().numpy()
    np.testing.assert_array_equal(g0, g1)
    np.testing.assert_array_equal(g1, got[:4])
    eng.close()
    ffma = CompiledNet(dlwp.model, N, force_ffma=True)
    assert not ffma.uses_tensor_cores()
    assert rel_err(ffma.predict(x0)[0], y.astype(np.float64)) < 5e-5
    ffma.close()


def test_unet_full_grid_on_tensor_cores(env):
    """BASELINE.json configs[2] shape (12, 180, 360): one application vs the torch-CPU tier-1 oracle."""
    nat, torch = env
    from dlwp_b200.engine import CompiledNet
    from tests.helpers import build_functional_pair
    cs = (12, 180, 360)
    dlwp, onet = build_functional_pair(cs, skip=True, integration_steps=1, seed=5)
    x0 = np.random.RandomState(6).standard_normal((2,) + cs).astype(np.float32)
    eng = CompiledNet(dlwp.model, 2, impl='tc')
    assert eng.uses_tensor_cores()
    y = eng.predict(x0)[0]
    assert nat.lib().dlwp_debug_flags() == 0
    with torch.no_grad():
        ref = onet.forward(torch.from_numpy(x0)).numpy()
    assert rel_err(y, ref.astype(np.float64)) < 5e-5
    eng.close()


# ---- magnitude sweeps (the unscaled split broke the bar for small inputs / small weights) ---------
def _check_scaled(nat, torch, cin, cout, k, d, act, xs, ws, bs, seed):
    rng = np.random.RandomState(seed)
    N, H, W = 2, 23, 60
    x = (xs * rng.standard_normal((N, cin, H, W))).astype(np.float32)
    w = (ws * OO.glorot_uniform(rng, k, k, cin, cout)).astype(np.float32)
    b = (bs * rng.standard_normal(cout)).astype(np.float32)
    pad = d * (k - 1) // 2
    pads = ((pad, pad), (pad, pad))
    y = run_conv(nat, torch, x, w, b, d, pads, nat.PAD_ZERO, nat.PAD_PERIODIC, act, nat.IMPL_TC)
    assert nat.lib().dlwp_debug_flags() == 0
    ref = OO.pad_conv2d_closed_form(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (d, d), pads[0],
                                    pads[1], 'zero', 'periodic')
    ref = OO.activation({0: None, 1: 'tanh', 2: 'relu'}[act])(ref)
    err = rel_err(y, ref)
    assert err < TOL, (xs, ws, err)


@pytest.mark.parametrize('xs', [1e-4, 1e-3, 1e-2, 1.0, 1e2, 1e4, 1e6])
@pytest.mark.parametrize('ws', [1e-3, 1e-2, 1.0, 10.0])
def test_layer_magnitude_sweep_conv2_geometry(env, xs, ws):
    nat, torch = env
    _check_scaled(nat, torch, 32, 6, 5, 1, nat.ACT_LINEAR, xs, ws, 0.0, 5)


@pytest.mark.parametrize('xs,ws', [(1e-4, 1.0), (1e-3, 1e-2), (1.0, 1e-3), (1e2, 1e-2), (1e4, 1e-4), (1e-4, 1e4)])
def test_layer_magnitude_sweep_conv1_geometry_tanh(env, xs, ws):
    """tanh layer: the products xs * ws keep the pre-activations O(1) or smaller.  (xs * ws >> 1 saturates tanh; an output
    near a zero crossing then carries the pre-activation's ABSOLUTE round-off, ~1e-7 * sum|w||x| -- ill-conditioned for
    any fp32 arithmetic, the FFMA kernels included -- so that regime cannot be held to a 2e-5 output bar.)"""
    nat, torch = env
    _check_scaled(nat, torch, 6, 32, 3, 2, nat.ACT_TANH, xs, ws, 0.1 * min(1.0, xs * ws), 6)


def test_trained_like_weights_per_layer(env):
    """weights ~ N(0, 0.01), bias ~ 0.1: what L2-regularised training (examples/train.py:155) produces."""
    nat, torch = env
    rng = np.random.RandomState(9)
    for cin, cout, k, d, act in ((6, 32, 3, 2, nat.ACT_TANH), (32, 6, 5, 1, nat.ACT_LINEAR)):
        x = rng.standard_normal((2, cin, 23, 60)).astype(np.float32)
        w = (0.01 * rng.standard_normal((k, k, cin, cout))).astype(np.float32)
        b = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        pad = d * (k - 1) // 2
        pads = ((pad, pad), (pad, pad))
        y = run_conv(nat, torch, x, w, b, d, pads, nat.PAD_ZERO, nat.PAD_PERIODIC, act, nat.IMPL_TC)
        assert nat.lib().dlwp_debug_flags() == 0
        ref = OO.pad_conv2d_closed_form(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (d, d), pads[0],
                                        pads[1], 'zero', 'periodic')
        ref = OO.activation({0: None, 1: 'tanh', 2: 'relu'}[act])(ref)
        assert rel_err(y, ref) < TOL


def _net_a_with_weights(shape, kernels, biases):
    from oracle import layers as OL
    from tests.helpers import build_product_sequential
    layers = OL.net_a_layers(shape)
    dlwp = build_product_sequential(layers)
    net = OL.OSequential(layers)
    ws = []
    for kk, bb in zip(kernels, biases):
        ws += [kk, bb]
    net.set_weights(ws)
    dlwp.model.set_weights(ws)
    return dlwp, net


@pytest.mark.parametrize('xs,w1s,w2s,bias', [(1e-4, 1.0, 1.0, 0.0), (1e-2, 1.0, 1.0, 0.02), (1e2, 1.0, 1.0, 0.02),
                                             (1e-4, 1e4, 1e-4, 0.0), (1e-2, 1e2, 1e-2, 0.0), (1e2, 1e-2, 1e2, 0.0),
                                             (1e4, 1e-4, 1e4, 0.0), (1.0, 1e-2, 1e-2, 0.1)])
@pytest.mark.parametrize('fuse', [0, 1])
def test_net_a_50_step_rollout_magnitude_sweep(env, xs, w1s, w2s, bias, fuse):
    """The BASELINE gate (1e-4 after 50 feedback steps) with scaled inputs / weights, on the tensor-core chain (fused or
    not), exponents of the state image decided on the device every step.  The cases keep the loop gain w1s * w2s at 1 (the
    state lives at magnitude w2s: same dynamics as the unscaled net) or let biases hold the state up; a gain >> 1 makes the
    rollout chaotic (any fp32 arithmetic diverges from float64 within 50 steps) and a gain << 1 without bias drives the
    state below fp32's own range -- neither can be held to a parity gate."""
    nat, torch = env
    from dlwp_b200.engine import CompiledNet
    from tests.helpers import oracle_rollout64
    shape = (6, 46, 92)
    rng = np.random.RandomState(21)
    k1 = (w1s * OO.glorot_uniform(rng, 3, 3, 6, 32)).astype(np.float32)
    k2 = (w2s * OO.glorot_uniform(rng, 5, 5, 32, 6)).astype(np.float32)
    b1 = (bias * rng.standard_normal(32)).astype(np.float32)
    b2 = (bias * rng.standard_normal(6)).astype(np.float32)
    dlwp, net = _net_a_with_weights(shape, (k1, k2), (b1, b2))
    x0 = (xs * rng.standard_normal((3,) + shape)).astype(np.float32)
    eng = CompiledNet(dlwp.model, 3, options={'fuse': fuse})   # fuse = 1: conv1 -> conv2 as one kernel (csrc/conv_fused.cu)
    assert eng.uses_tensor_cores() and (eng.fused_pair() == 0) == bool(fuse)
    got = eng.rollout_device(torch.from_numpy(x0).cuda(), 50, use_graph=True).cpu().numpy()
    assert nat.lib().dlwp_debug_flags() == 0
    ref = oracle_rollout64(net, x0, 50)
    per_step = [rel_err(got[t], ref[t]) for t in range(50)]
    bad = np.where(np.abs(got[0] - ref[0]) > 1e-4 * np.abs(ref[0]).max())
    assert max(per_step) <= 1e-4, (max(per_step), per_step[:3], 'step 0: %d bad elements, samples %s rows %s cols %s' % (
        len(bad[0]), np.unique(bad[0]), np.unique(bad[2])[:12], np.unique(bad[3])[:12]), eng.fused_pair())
    assert eng.uses_tensor_cores()
    eng.close()


def test_trained_like_net_a_rollout(env):
    """weights ~ N(0, 0.01) (conv2: N(0, 0.03) so the state does not die out), bias ~ 0.1, 50 steps, 1e-4."""
    nat, torch = env
    from dlwp_b200.engine import CompiledNet
    from tests.helpers import oracle_rollout64
    shape = (6, 46, 92)
    rng = np.random.RandomState(22)
    k1 = (0.01 * rng.standard_normal((3, 3, 6, 32))).astype(np.float32)
    k2 = (0.03 * rng.standard_normal((5, 5, 32, 6))).astype(np.float32)
    b1 = (0.1 * rng.standard_normal(32)).astype(np.float32)
    b2 = (0.1 * rng.standard_normal(6)).astype(np.float32)
    dlwp, net = _net_a_with_weights(shape, (k1, k2), (b1, b2))
    x0 = rng.standard_normal((3,) + shape).astype(np.float32)
    eng = CompiledNet(dlwp.model, 3)
    assert eng.uses_tensor_cores()
    got = eng.rollout_host(x0, 50)
    assert eng.uses_tensor_cores()                         # no fallback happened
    ref = oracle_rollout64(net, x0, 50)
    assert max(rel_err(got[t], ref[t]) for t in range(50)) <= 1e-4
    eng.close()


def test_tanh_underflow_is_flagged_and_falls_back(env):
    """A tanh layer whose outputs are ALL ~1e-9 underflows the static 2^14 scale of its image: flagged, rerun on FFMA."""
    nat, torch = env
    import warnings
    from dlwp_b200.engine import CompiledNet
    from tests.helpers import oracle_rollout64
    shape = (6, 20, 36)
    rng = np.random.RandomState(23)
    k1 = (1e-9 * OO.glorot_uniform(rng, 3, 3, 6, 32)).astype(np.float32)
    k2 = OO.glorot_uniform(rng, 5, 5, 32, 6)
    dlwp, net = _net_a_with_weights(shape, (k1, k2), (np.zeros(32, np.float32), np.zeros(6, np.float32)))
    x0 = rng.standard_normal((2,) + shape).astype(np.float32)
    eng = CompiledNet(dlwp.model, 2)
    assert eng.uses_tensor_cores()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        y = eng.predict(x0)[0]
    assert any('underflow' in str(m.message) for m in w)
    assert not eng.uses_tensor_cores()
    assert rel_err(y, net.forward(x0.astype(np.float64))) < 2e-5
    eng.close()


# ---- ADVICE r01 (plan.cu tc_setup) -------------------------------------------------------------------------------------
def test_single_conv_same_channels_does_not_feed_back_into_its_own_input(env):
    """One conv with Cin == Cout: feeding the output back into the input image inside the same launch would let CTAs
    overwrite rows other CTAs still read.  The plan must pack the state instead; rollout == oracle, graph == no graph."""
    nat, torch = env
    from dlwp_b200.engine import CompiledNet
    from oracle import layers as OL
    from tests.helpers import build_product_sequential, oracle_rollout64, oracle_sequential_like
    cf = 'channels_first'
    shape = (8, 40, 96)
    layers = (('PeriodicPadding2D', ((0, 1),), {'data_format': cf, 'input_shape': shape}),
              ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
              ('Conv2D', (8, 3), {'padding': 'valid', 'activation': 'tanh', 'data_format': cf}))
    dlwp = build_product_sequential(layers)
    net = oracle_sequential_like(dlwp, layers, seed=4, bias_scale=0.05)
    x0 = np.random.RandomState(5).standard_normal((5,) + shape).astype(np.float32)
    eng = CompiledNet(dlwp.model, 5, impl='tc')
    assert eng.uses_tensor_cores()
    xd = torch.from_numpy(x0).cuda()
    g = eng.rollout_device(xd, 6, use_graph=True).cpu().numpy()
    p = eng.rollout_device(xd, 6, use_graph=False).cpu().numpy()
    np.testing.assert_array_equal(g, p)
    ref = oracle_rollout64(net, x0, 6)
    assert rel_err(g, ref) < 2e-5
    np.testing.assert_array_equal(eng.predict(x0)[0], g[0])
    eng.close()


def test_model_ending_in_a_data_mover_stays_off_the_p_only_path(env):
    """UpSampling2D / an aligned slice as the model OUTPUT: the P-image data movers have no fp32 destination, so such a plan
    must not run as a tensor-core chain (it returned uninitialised memory in round 1)."""
    nat, torch = env
    from dlwp_b200 import keras
    from dlwp_b200.custom import PeriodicPadding2D, slice_layer
    from dlwp_b200.engine import CompiledNet
    from dlwp_b200.keras.layers import Conv2D, Input, UpSampling2D, ZeroPadding2D
    cf = 'channels_first'
    x_in = Input(shape=(8, 12, 40))
    conv = Conv2D(16, 3, padding='valid', activation='tanh', data_format=cf)
    t = conv(PeriodicPadding2D(padding=(0, 1), data_format=cf)(ZeroPadding2D(padding=(1, 0), data_format=cf)(x_in)))
    rng = np.random.RandomState(8)
    k = OO.glorot_uniform(rng, 3, 3, 8, 16)
    b = (0.05 * rng.standard_normal(16)).astype(np.float32)
    x = rng.standard_normal((2, 8, 12, 40)).astype(np.float32)
    ref = np.tanh(OO.pad_conv2d_closed_form(x.astype(np.float64), k.astype(np.float64), b.astype(np.float64), (1, 1),
                                            (1, 1), (1, 1), 'zero', 'periodic'))
    for outputs, want in ((UpSampling2D(2, data_format=cf)(t), np.repeat(np.repeat(ref, 2, axis=2), 2, axis=3)),
                          (slice_layer(8, 16, axis=1)(t), ref[:, 8:16]),
                          ([t, t], None)):
        model = keras.Model(inputs=x_in, outputs=outputs)
        conv.set_weights([k, b])
        eng = CompiledNet(model, 2)
        ys = eng.predict(x)
        if want is None:
            assert rel_err(ys[0], ref) < 2e-5 and rel_err(ys[1], ref) < 2e-5
        else:
            assert ys[0].shape == want.shape and rel_err(ys[0], want) < 2e-5
        eng.close()
