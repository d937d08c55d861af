"""
Latitude-band partitioning: band planner arithmetic, and a world_size-2/3 gloo run of the halo exchange driver on CPU.
The band forward used here is the ORACLE (float64) evaluated on a state whose rows outside the planner's `need_in` window
are poisoned with NaN -- so the test fails if the planner under-estimates the halo or the exchange delivers wrong rows.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import layers as OL
from oracle import rollout as OR


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _net_a(cs):
    from dlwp_b200.engine import Lowering
    from tests.helpers import build_product_sequential, oracle_sequential_like
    layers = OL.net_a_layers(cs)
    dlwp = build_product_sequential(layers)
    net = oracle_sequential_like(dlwp, layers, seed=1, bias_scale=0.05)
    return Lowering(dlwp.model), net


def test_band_bounds_and_net_a_halo():
    from dlwp_b200.parallel import BandPlanner, band_bounds, make_planners
    assert band_bounds(91, 8) == [(0, 12), (12, 24), (24, 36), (36, 47), (47, 58), (58, 69), (69, 80), (80, 91)]
    with pytest.raises(ValueError):
        band_bounds(4, 5)
    low, _ = _net_a((6, 91, 180))
    planners = make_planners(low.ops, low.buffers, 91, 8)
    assert planners[0].halo == (0, 4) and planners[7].halo == (4, 0)       # poles: zero padding, not halo
    assert all(p.halo == (4, 4) for p in planners[1:7])                    # 2 (conv1, d=2) + 2 (conv2, k=5)
    p = planners[3]
    assert p.windows[1] == p.band                                          # last conv: exactly the band
    assert p.windows[0] == (p.band[0] - 2, p.band[1] + 2)                  # first conv: band +- conv2's reach
    assert 1.0 < p.redundancy(low.ops, low.buffers) < 1.2
    one = BandPlanner(low.ops, low.buffers, (0, 91))
    assert one.halo == (0, 0) and one.windows == [(0, 91), (0, 91)]


def test_unet_halo_is_fourteen_rows():
    from dlwp_b200.engine import Lowering
    from dlwp_b200.parallel import make_planners
    from tests.helpers import build_functional_pair
    dlwp, _ = build_functional_pair((12, 64, 32), skip=True, integration_steps=1)
    low = Lowering(dlwp.model)
    planners = make_planners(low.ops, low.buffers, 64, 2)
    # 2 + 2*1 + 4*1 (three levels down) + 2*1 + 2 (up) + 2 -> 14 rows, rounded outward by the pool / upsample alignment
    assert 14 <= planners[0].halo[1] <= 16 and 14 <= planners[1].halo[0] <= 16


def _worker(rank, world, port, cs, n, steps, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from dlwp_b200.parallel import LatBandRollout, make_planners
        low, net = _net_a(cs)
        H = cs[1]
        planners = make_planners(low.ops, low.buffers, H, world)
        me = planners[rank]

        def forward(src, outs):
            x = src.numpy().astype(np.float64).copy()
            lo, hi = me.need_in
            x[:, :, :lo] = np.nan                      # rows this rank has no right to read
            x[:, :, hi:] = np.nan
            xin = np.nan_to_num(x, nan=1e30)           # poison that survives tanh as a huge, obviously wrong value
            y = net.forward(xin)
            outs[0][:, :, me.band[0]:me.band[1]] = torch.from_numpy(y[:, :, me.band[0]:me.band[1]].astype(np.float32))

        x0 = torch.from_numpy(np.random.RandomState(0).standard_normal((n,) + cs).astype(np.float32))
        series = torch.full((steps, n) + cs, float('nan'))
        drv = LatBandRollout(H, rank, world, planners, forward, dist=dist)
        drv.rollout(x0, series, steps)
        band = series[:, :, :, me.band[0]:me.band[1]].contiguous()
        gathered = [torch.empty((steps, n, cs[0], b.band[1] - b.band[0], cs[2])) for b in planners]
        dist.all_gather(gathered, band) if len({g.shape for g in gathered}) == 1 else _gather_uneven(gathered, band, rank, world)
        if rank == 0:
            full = torch.cat(gathered, dim=3).numpy()
            ref = OR.neuralnet_predict_timeseries(lambda p: net.forward(p), x0.numpy().astype(np.float64), steps,
                                                  dtype=np.float64)
            ret['err'] = float(np.abs(full - ref).max() / np.abs(ref).max())
            ret['halo_bytes'] = drv.halo_bytes_per_iteration
    finally:
        dist.destroy_process_group()


def _gather_uneven(gathered, band, rank, world):
    for r in range(world):
        t = band if r == rank else gathered[r]
        dist.broadcast(t, src=r)
        gathered[r].copy_(t) if r != rank else gathered[r].copy_(band)


def _simulate_row_validity(low, planner, nat):
    """Execute the plan symbolically for one rank: valid[buffer][channel, row] starts true on the input rows the planner asks
    for (`need_in`) and becomes true where an op writes; every destination row an op computes must find ALL the source rows
    its definition reads (pad / dilation / pool / upsample maps written out per row, not the planner's interval arithmetic)
    valid or outside [0, H) (latitude padding).  Returns valid[] at the end."""
    bufs = low.buffers
    valid = [np.zeros((b['C'], b['H']), bool) for b in bufs]
    in_buf = [i for i, b in enumerate(bufs) if b['kind'] == nat.BUF_INPUT][0]
    valid[in_buf][:, planner.need_in[0]:planner.need_in[1]] = True
    for op, win in zip(low.ops, planner.windows):
        if win is None:
            continue
        src, dst = valid[op['src']], valid[op['dst']]
        Hs = bufs[op['src']]['H']
        ch = slice(op['src_c0'], op['src_c0'] + op['src_c'])
        kind = op['kind']

        def rows_read(y):
            if kind == nat.OP_CONV:
                rows = [y - op['pad_t'] + op['dil_h'] * i for i in range(op['kh'])]     # rows of the (pre-op'ed) source
                Hp = Hs // 2 if op['pre_op'] == 1 else (Hs * 2 if op['pre_op'] == 2 else Hs)
                rows = [r for r in rows if 0 <= r < Hp]                                  # others are zero padding
                if op['pre_op'] == 1:
                    return [q for r in rows for q in (2 * r, 2 * r + 1)]
                if op['pre_op'] == 2:
                    return [r // 2 for r in rows]
                return rows
            if kind == nat.OP_PAD:
                return [r for r in [y - op['pad_t']] if 0 <= r < Hs]
            if kind == nat.OP_MAXPOOL:
                return [2 * y, 2 * y + 1]
            if kind == nat.OP_UPSAMPLE:
                return [y // 2]
            return [y]

        for y in range(win[0], win[1]):
            for r in rows_read(y):
                assert src[ch, r].all(), (op, win, y, r)
        out_c = op['Cout'] if kind == nat.OP_CONV else op['src_c']
        dst[op['dst_c0']:op['dst_c0'] + out_c, win[0]:win[1]] = True
    return valid


@pytest.mark.parametrize('world', [1, 2, 3, 4, 5, 8])
def test_band_windows_cover_every_row_their_consumers_read(world):
    """Net A (91 rows) and the skip U-Net (two 2x poolings, upsamplings, skip concatenations; one and two unrolled
    applications): for every rank the row windows suffice, the outputs hold the whole band, and the bands tile [0, H)."""
    from dlwp_b200 import _native as nat
    from dlwp_b200.engine import Lowering
    from dlwp_b200.parallel import make_planners
    from tests.helpers import build_functional_pair
    cases = [(_net_a((6, 91, 180))[0], 91)]
    for steps in (1, 2):
        dlwp, _ = build_functional_pair((12, 180, 36), skip=True, integration_steps=steps)
        cases.append((Lowering(dlwp.model), 180))
    for low, H in cases:
        planners = make_planners(low.ops, low.buffers, H, world)
        assert [p.band for p in planners][0][0] == 0 and planners[-1].band[1] == H
        assert all(a.band[1] == b.band[0] for a, b in zip(planners, planners[1:]))
        for p in planners:
            valid = _simulate_row_validity(low, p, nat)
            for i, b in enumerate(low.buffers):
                if b['kind'] == nat.BUF_OUTPUT:
                    assert valid[i][:, p.band[0]:p.band[1]].all()
            assert p.halo[0] >= 0 and p.halo[1] >= 0


@pytest.mark.parametrize('world', [2, 3])
def test_latband_rollout_with_halo_exchange_matches_single_domain(world):
    cs, n, steps = (6, 23, 16), 2, 4
    port = _free_port()
    mgr = mp.get_context('spawn').Manager()     # (never fork a process that already runs helper threads)
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, cs, n, steps, ret), nprocs=world, join=True)
    assert ret['err'] < 1e-6, ret['err']                 # float32 storage of float64 results between iterations
    assert ret['halo_bytes'] == (4 if world == 2 else 4) * n * cs[0] * cs[2] * 4
