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


def test_out_of_range_inputs_fall_back_to_fp32_kernels(env):
    """|x| > 65504 cannot be split into fp16 hi/lo: the kernels flag it and the engine reruns on the FFMA path."""
    nat, torch = env
    import warnings
    from dlwp_b200.engine import CompiledNet
    from oracle import layers as OL
    from tests.helpers import build_product_sequential, oracle_sequential_like
    layers = OL.net_a_layers((6, 20, 36))
    dlwp = build_product_sequential(layers)
    net = oracle_sequential_like(dlwp, layers, seed=2, bias_scale=0.0)
    x0 = np.random.RandomState(1).standard_normal((2, 6, 20, 36)).astype(np.float32)
    x0[0, 0, 3, 5] = 3.0e5
    eng = CompiledNet(dlwp.model, 2)
    assert eng.uses_tensor_cores()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        y = eng.predict(x0)[0]
    assert any('fp16-split range' in str(m.message) for m in w)
    assert not eng.uses_tensor_cores()
    assert rel_err(y, net.forward(x0.astype(np.float64))) < 2e-5
    eng.close()


def test_latitude_band_windows_on_tensor_cores(env):
    """Row-windowed tensor-core plans (one per band) reproduce the single-domain tensor-core rollout bit for bit."""
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
    np.testing.assert_array_equal(got, ref)


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


@pytest.mark.parametrize('shape,N', [((6, 91, 180), 3), ((6, 20, 36), 2)])
def test_fp32_state_first_layer_matches_the_p_layout_path_bit_for_bit(env, monkeypatch, shape, N):
    """DLWP_SW_F32IN=1: the first layer stages raw fp32 rows of the state and converter warps build the hi/lo A layout
    (periodic wrap and pole rows included); no P image of the state, no pack kernel, no feedback copy.  The split of a value
    is the same either way, so the rollout is bit-identical to the default tensor-core chain."""
    nat, torch = env
    from dlwp_b200.engine import CompiledNet
    from oracle import layers as OL
    from tests.helpers import build_product_sequential, oracle_sequential_like
    layers = OL.net_a_layers(shape)
    dlwp = build_product_sequential(layers)
    oracle_sequential_like(dlwp, layers, seed=1, bias_scale=0.02)
    x0 = np.random.RandomState(0).standard_normal((N,) + shape).astype(np.float32)
    xd = torch.from_numpy(x0).cuda()
    eng = CompiledNet(dlwp.model, N, impl='tc')
    ref = eng.rollout_device(xd, 6, use_graph=True).cpu().numpy()
    eng.close()
    monkeypatch.setenv('DLWP_SW_F32IN', '1')
    eng2 = CompiledNet(dlwp.model, N, impl='tc')
    assert eng2.uses_tensor_cores()
    got = eng2.rollout_device(xd, 6, use_graph=False).cpu().numpy()
    assert nat.lib().dlwp_debug_flags() == 0
    np.testing.assert_array_equal(got, ref)
    got_g = eng2.rollout_device(xd, 6, use_graph=True).cpu().numpy()
    np.testing.assert_array_equal(got_g, ref)
    eng2.close()
