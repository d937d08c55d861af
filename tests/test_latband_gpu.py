"""
Latitude-band execution on the GPU: P row-windowed plans run in ONE process (the exchange is emulated with device copies
between the per-rank series), and the assembled bands must equal the single-domain rollout to fp32 round-off -- same
kernels, same per-pixel summation order, only the tile origins differ.  (Not bit for bit any more: this driver re-packs the
state from a series slot every iteration, and the exponent of the scaled fp16 split then comes from the amax of the rows
the band reads, which differs from the global amax; the native rollout (dlwp_rollout_latband) derives the state's exponent
from the weights alone and stays bit-identical -- bench.py checks that on every multi-GPU run.)
"""

ROUND_OFF = 1e-6   # of max|ref|

import numpy as np
import pytest

from oracle import layers as OL
from tests.helpers import build_functional_pair, build_product_sequential, oracle_sequential_like

pytestmark = pytest.mark.gpu


class _FakeDist(object):
    """Single-process stand-in for torch.distributed P2P between emulated ranks (mailbox keyed by (src, dst))."""

    class P2POp(object):
        def __init__(self, op, tensor, peer, group=None):
            self.op, self.tensor, self.peer = op, tensor, peer

    isend, irecv = 'isend', 'irecv'


def _run_bands(model, world, x0, iterations):
    import torch
    from dlwp_b200.engine import CompiledNet, Lowering
    from dlwp_b200.parallel import make_planners
    low = Lowering(model)
    H = x0.shape[2]
    planners = make_planners(low.ops, low.buffers, H, world)
    nets = [CompiledNet(model, x0.shape[0], row_windows=p.windows) for p in planners]
    n_out = nets[0].n_outputs
    xd = torch.from_numpy(x0).cuda()
    series = [torch.full((iterations * n_out,) + x0.shape, float('nan'), device='cuda') for _ in range(world)]
    for t in range(iterations):
        for r in range(world):
            src = xd if t == 0 else series[r][t * n_out - 1]
            nets[r].forward_into(src, [series[r][t * n_out + k] for k in range(n_out)])
        last = t * n_out + n_out - 1
        for r in range(world):                      # emulated halo exchange of the last output
            lo, hi = planners[r].band
            top, bot = planners[r].halo
            if top:
                series[r][last][:, :, lo - top:lo] = series[r - 1][last][:, :, lo - top:lo]
            if bot:
                series[r][last][:, :, hi:hi + bot] = series[r + 1][last][:, :, hi:hi + bot]
    full = torch.cat([series[r][:, :, :, p.band[0]:p.band[1]] for r, p in enumerate(planners)], dim=3)
    for n in nets:
        n.close()
    return full.cpu().numpy(), planners


@pytest.mark.parametrize('world', [2, 8])
def test_net_a_bands_equal_single_domain_to_round_off(world):
    layers = OL.net_a_layers()
    dlwp = build_product_sequential(layers)
    oracle_sequential_like(dlwp, layers, seed=1, bias_scale=0.05)
    x0 = np.random.RandomState(0).standard_normal((3, 6, 91, 180)).astype(np.float32)
    ref = dlwp.predict_timeseries(x0, 6)
    got, planners = _run_bands(dlwp.model, world, x0, 6)
    assert np.isfinite(got).all() and np.abs(got - ref).max() <= ROUND_OFF * np.abs(ref).max()
    assert planners[1].halo[0] == 4


def test_unet_bands_equal_single_domain():
    cs = (12, 48, 64)
    dlwp, _ = build_functional_pair(cs, skip=True, integration_steps=2, seed=3)
    x0 = np.random.RandomState(4).standard_normal((2,) + cs).astype(np.float32)
    # Row-windowed U-Net plans run as tensor-core chains too (row windows in the convs and in the P-image data movers);
    # every pixel sums its taps in the same order whatever the band, so the bands reproduce the single-domain chain bit
    # for bit, and both agree with the fp32 plan to round-off.
    from dlwp_b200.engine import CompiledNet
    assert dlwp.model.engine(2).uses_tensor_cores()
    ref = dlwp.predict_timeseries(x0, 4)
    got, planners = _run_bands(dlwp.model, 2, x0, 2)
    assert np.isfinite(got).all() and np.abs(got - ref).max() <= ROUND_OFF * np.abs(ref).max()
    eng = CompiledNet(dlwp.model, 2, force_ffma=True)
    fp32 = eng.rollout_host(x0, 2)
    eng.close()
    assert fp32.shape == ref.shape
    assert np.abs(fp32 - ref).max() <= 5e-5 * np.abs(ref).max()


def _run_bands_p2p(model, world, x0, iterations, use_graph, options=None):
    """The NATIVE latitude-band rollout (dlwp_rollout_latband) with the halo over peer memory, `world` bands in one process:
    one row-windowed plan per band, linked by raw pointers (dlwp_plan_halo_connect), each launched on its own stream -- the
    arrival counters synchronise the streams exactly as they synchronise GPUs."""
    import ctypes
    import torch
    from dlwp_b200 import _native as nat
    from dlwp_b200.engine import CompiledNet, Lowering
    from dlwp_b200.parallel import make_planners
    lib = nat.lib()
    low = Lowering(model)
    planners = make_planners(low.ops, low.buffers, x0.shape[2], world)
    nets = [CompiledNet(model, x0.shape[0], row_windows=p.windows, options=options) for p in planners]
    for n in nets:
        nat.check(lib.dlwp_plan_halo_enable(n.plan), 'dlwp_plan_halo_enable')
    for r in range(world):
        if r > 0:
            nat.check(lib.dlwp_plan_halo_connect(nets[r].plan, 0, nets[r - 1].plan))
        if r + 1 < world:
            nat.check(lib.dlwp_plan_halo_connect(nets[r].plan, 1, nets[r + 1].plan))
    n_out = nets[0].n_outputs
    xd = torch.from_numpy(x0).cuda()
    series = [torch.full((iterations * n_out,) + x0.shape, float('nan'), device='cuda') for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for rep in range(2):                                    # twice: counters and image parities carry over between rollouts
        for r in range(world):
            p = planners[r]
            up = planners[r - 1] if r > 0 else None
            down = planners[r + 1] if r + 1 < world else None
            info = nat.BandInfo(r, world, p.band[0], p.band[1], p.halo[0], p.halo[1], up.halo[1] if up else 0,
                                down.halo[0] if down else 0)
            nets[r].sync_weights()
            nat.check(lib.dlwp_rollout_latband(nets[r].plan, None, x0.shape[0], xd.data_ptr(), series[r].data_ptr(),
                                               iterations, ctypes.byref(info), 1 if use_graph else 0,
                                               ctypes.c_void_p(streams[r].cuda_stream)), 'dlwp_rollout_latband')
        torch.cuda.synchronize()
        assert lib.dlwp_debug_flags() == 0                  # (bit 0 would be a halo wait that timed out)
    full = torch.cat([series[r][:, :, :, p.band[0]:p.band[1]] for r, p in enumerate(planners)], dim=3)
    for n in nets:
        n.close()
    return full.cpu().numpy()


@pytest.mark.parametrize('world,use_graph', [(2, False), (3, True), (4, True)])
def test_net_a_peer_memory_halo_is_bit_identical_to_single_domain(world, use_graph):
    layers = OL.net_a_layers()
    dlwp = build_product_sequential(layers)
    oracle_sequential_like(dlwp, layers, seed=1, bias_scale=0.05)
    x0 = np.random.RandomState(0).standard_normal((3, 6, 91, 180)).astype(np.float32)
    ref = dlwp.predict_timeseries(x0, 7)
    got = _run_bands_p2p(dlwp.model, world, x0, 7, use_graph)
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_unet_peer_memory_halo_is_bit_identical_to_single_domain(precision):
    import torch
    from dlwp_b200.engine import CompiledNet
    cs = (12, 48, 64)
    dlwp, _ = build_functional_pair(cs, skip=True, integration_steps=1, seed=3)
    x0 = np.random.RandomState(4).standard_normal((2,) + cs).astype(np.float32)
    eng = CompiledNet(dlwp.model, 2, options={'precision': precision})
    ref = eng.rollout_device(torch.from_numpy(x0).cuda(), 4, use_graph=False).cpu().numpy()
    eng.close()
    got = _run_bands_p2p(dlwp.model, 2, x0, 4, True, options={'precision': precision})
    np.testing.assert_array_equal(got, ref)


def test_latband_host_rollout_pipelined_band_copies():
    """dlwp_rollout_latband_host (numpy in, this band's rows of every state out; groups of steps copied D2H while the next
    ones compute): three bands in one process, one blocking call per band on its own thread, peer-memory halo -- the
    concatenated bands equal the single-domain predict_timeseries bit for bit, twice in a row (counters carry over)."""
    import ctypes
    import threading
    import torch
    from dlwp_b200 import _native as nat
    from dlwp_b200.engine import CompiledNet, Lowering
    from dlwp_b200.parallel import make_planners
    lib = nat.lib()
    layers = OL.net_a_layers()
    dlwp = build_product_sequential(layers)
    oracle_sequential_like(dlwp, layers, seed=5, bias_scale=0.05)
    x0 = np.random.RandomState(2).standard_normal((3, 6, 91, 180)).astype(np.float32)
    iterations, world = 7, 3
    ref = dlwp.predict_timeseries(x0, iterations)
    low = Lowering(dlwp.model)
    planners = make_planners(low.ops, low.buffers, 91, world)
    nets = [CompiledNet(dlwp.model, 3, row_windows=p.windows) for p in planners]
    for n in nets:
        nat.check(lib.dlwp_plan_halo_enable(n.plan), 'dlwp_plan_halo_enable')
        n.sync_weights()
    for r in range(world):
        if r > 0:
            nat.check(lib.dlwp_plan_halo_connect(nets[r].plan, 0, nets[r - 1].plan))
        if r + 1 < world:
            nat.check(lib.dlwp_plan_halo_connect(nets[r].plan, 1, nets[r + 1].plan))
    for rep in range(2):
        bands = [torch.full((iterations, 3, 6, p.band[1] - p.band[0], 180), float('nan'), dtype=torch.float32,
                            pin_memory=True) for p in planners]
        codes = [None] * world

        def work(r):
            p = planners[r]
            up = planners[r - 1] if r > 0 else None
            down = planners[r + 1] if r + 1 < world else None
            info = nat.BandInfo(r, world, p.band[0], p.band[1], p.halo[0], p.halo[1], up.halo[1] if up else 0,
                                down.halo[0] if down else 0)
            codes[r] = lib.dlwp_rollout_latband_host(nets[r].plan, None, 3, x0.ctypes.data, bands[r].data_ptr(),
                                                     iterations, ctypes.byref(info), 2)

        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert codes == [0] * world, (codes, lib.dlwp_last_error_string())
        assert lib.dlwp_debug_flags() == 0
        got = np.concatenate([b.numpy() for b in bands], axis=3)
        np.testing.assert_array_equal(got, ref)
    for n in nets:
        n.close()


def test_latband_engine_predict_timeseries_single_rank():
    """LatBandEngine.predict_timeseries (the reference-facing entry of the lat-band path) with one rank: equals the
    single-domain predict_timeseries bit for bit; multi-rank runs are checked by bench.py / scripts/latband_check.py."""
    from dlwp_b200.parallel import LatBandEngine
    layers = OL.net_a_layers((6, 30, 60))
    dlwp = build_product_sequential(layers)
    oracle_sequential_like(dlwp, layers, seed=2, bias_scale=0.05)
    x0 = np.random.RandomState(1).standard_normal((3, 6, 30, 60)).astype(np.float32)
    eng = LatBandEngine(dlwp.model, 3, 0, 1)
    got = eng.predict_timeseries(x0, 5)
    np.testing.assert_array_equal(got, dlwp.predict_timeseries(x0, 5))
    with pytest.raises(ValueError):
        eng.predict_timeseries(x0, 0)
    eng.close()
