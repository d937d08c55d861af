"""
Latitude-band domain decomposition of the rollout across the GPUs of one box: one process per GPU, one grouped
SendRecv of halo rows per rollout iteration.

The reference has no spatial parallelism at all (its only multi-device mechanism is keras `multi_gpu_model`, a
single-process batch split: DLWP/model/models.py:104-109); this module is new.  Longitude stays whole on every rank so
the periodic wrap (PeriodicPadding2D, DLWP/custom.py:191-214) is rank-local; latitude (H) is split into P contiguous
bands.  Rows beyond the poles are the ZeroPadding2D zeros of every layer, never halo data.

How a rank runs one iteration
-----------------------------
Every buffer keeps its full (N, C, H, W) shape, but each op only computes the destination rows that the rank's band
of the model outputs transitively needs (`BandPlanner`: a backward sweep over the lowered op list; a conv needs
`[y0 - pad_t, y1 - pad_t + dil*(kh-1))` of its source, a pool `[2*y0, 2*y1)`, an upsample `[y0//2, ceil(y1/2))`).  The
rows of the *input* state outside the band that this sweep reaches are the halo: `top = band_lo - need_lo`,
`bottom = need_hi - band_hi`.  After an iteration, the band rows of the last output (the next input) adjacent to the
neighbours are exchanged: one `batch_isend_irecv` (a single NCCL group: ncclSend/ncclRecv to/from both neighbours) per
iteration.  Net A: halo = 2 + 2 = 4 rows -> 4*180*6*4 B = 17,280 B per neighbour, direction and sample; the U-Net
(two 2x poolings, 6 convs) 14 rows.

The exchange logic is backend-agnostic (`torch.distributed` with NCCL on GPUs, gloo in the CPU tests); the band
forward is supplied by the caller (GPU: a row-windowed DlwpPlan; CPU tests: the oracle).
"""

import math

import numpy as np

from . import _native as nat


def band_bounds(H, parts):
    """Split H rows into `parts` contiguous bands whose sizes differ by at most one row (north to south)."""
    if parts < 1 or parts > H:
        raise ValueError('cannot split %d rows into %d bands' % (H, parts))
    base, extra = divmod(H, parts)
    bounds, lo = [], 0
    for r in range(parts):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


class BandPlanner(object):
    """
    Row windows for one rank.  `ops` / `buffers` are the dict lists of engine.Lowering; `band` = (lo, hi) rows of every
    model OUTPUT that this rank owns.  After construction:
      windows[i]   destination rows (lo, hi) op i must compute (None: op not needed at all)
      need_in      rows (lo, hi) of the input state that must be valid before an iteration
      halo         (top, bottom) = (band_lo - need_in_lo, need_in_hi - band_hi)
    """

    def __init__(self, ops, buffers, band):
        self.band = tuple(band)
        need = {}

        def add(buf, lo, hi):
            H = buffers[buf]['H']
            lo, hi = max(0, lo), min(H, hi)
            if hi <= lo:
                return
            if buf in need:
                need[buf] = (min(need[buf][0], lo), max(need[buf][1], hi))
            else:
                need[buf] = (lo, hi)

        in_buf = None
        for i, b in enumerate(buffers):
            if b['kind'] == nat.BUF_OUTPUT:
                add(i, band[0], band[1])
            if b['kind'] == nat.BUF_INPUT:
                in_buf = i
        self.windows = [None] * len(ops)
        for i in range(len(ops) - 1, -1, -1):
            op = ops[i]
            if op['dst'] not in need:
                continue
            lo, hi = need[op['dst']]
            self.windows[i] = (lo, hi)
            kind = op['kind']
            if kind == nat.OP_CONV:
                if op['pad_mode_h'] == nat.PAD_PERIODIC and (op['pad_t'] or op['pad_b']):
                    raise NotImplementedError('latitude bands need non-periodic latitude padding')
                slo = lo - op['pad_t']
                shi = hi - op['pad_t'] + op['dil_h'] * (op['kh'] - 1)
                if op['pre_op'] == 1:
                    slo, shi = 2 * slo, 2 * shi
                elif op['pre_op'] == 2:
                    slo, shi = slo // 2, -(-shi // 2)
            elif kind == nat.OP_PAD:
                if op['pad_mode_h'] != nat.PAD_ZERO and (op['pad_t'] or op['pad_b']):
                    raise NotImplementedError('latitude bands need zero latitude padding')
                slo, shi = lo - op['pad_t'], hi - op['pad_t']
            elif kind == nat.OP_MAXPOOL:
                slo, shi = 2 * lo, 2 * hi
            elif kind == nat.OP_UPSAMPLE:
                slo, shi = lo // 2, -(-hi // 2)
            else:
                slo, shi = lo, hi
            add(op['src'], slo, shi)
        if in_buf is None or in_buf not in need:
            raise ValueError('the model outputs do not depend on the input')
        self.need_in = need[in_buf]
        self.halo = (band[0] - self.need_in[0], self.need_in[1] - band[1])
        if min(self.halo) < 0:
            raise ValueError('band %s is not covered by the rows the net reads %s' % (band, self.need_in))
        self.need = need

    def redundancy(self, ops, buffers):
        """Rows computed by this rank / rows it would compute with an exact partition (conv ops only)."""
        done = exact = 0.0
        for op, w in zip(ops, self.windows):
            if w is None or op['kind'] != nat.OP_CONV:
                continue
            H = buffers[op['dst']]['H']
            scale = H / float(buffers[[i for i, b in enumerate(buffers) if b['kind'] == nat.BUF_OUTPUT][0]]['H'])
            work = op['kh'] * op['kw'] * op['src_c'] * op['Cout'] * buffers[op['dst']]['W']
            done += work * (w[1] - w[0])
            exact += work * (self.band[1] - self.band[0]) * scale
        return done / max(exact, 1.0)


class LatBandRollout(object):
    """
    The per-rank rollout driver.  `forward(src, outs)` must compute, for full-shape tensors, at least the band rows of
    every output from the rows `need_in` of `src` (the GPU implementation is `CompiledNet.forward_into`; the CPU tests
    pass an oracle-based callable).  `dist` is `torch.distributed` (or None for a single band).
    """

    def __init__(self, H, rank, world, planners, forward, dist=None, group=None):
        self.H, self.rank, self.world = H, rank, world
        self.planners = planners            # one BandPlanner per rank (halo sizes of the neighbours are needed too)
        self.me = planners[rank]
        self.forward = forward
        self.dist = dist
        self.group = group
        self.halo_bytes_per_iteration = 0
        for r, p in enumerate(planners):
            lo, hi = p.band
            if r > 0 and p.halo[0] > planners[r - 1].band[1] - planners[r - 1].band[0]:
                raise ValueError('halo of %d rows exceeds the neighbouring band' % p.halo[0])
            if r < world - 1 and p.halo[1] > planners[r + 1].band[1] - planners[r + 1].band[0]:
                raise ValueError('halo of %d rows exceeds the neighbouring band' % p.halo[1])

    def exchange(self, slot):
        """Fill the halo rows of `slot` (N, C, H, W) from the neighbours and serve theirs: ONE grouped SendRecv."""
        if self.world == 1:
            return 0
        import torch
        dist = self.dist
        lo, hi = self.me.band
        ops, recvs, nbytes = [], [], 0
        up, down = self.rank - 1, self.rank + 1
        if up >= 0:
            k = self.planners[up].halo[1]            # rows the upper neighbour needs from the top of my band
            if k:
                send = slot[:, :, lo:lo + k, :].contiguous()
                ops.append(dist.P2POp(dist.isend, send, up, self.group))
                nbytes += send.numel() * 4
            k = self.me.halo[0]
            if k:
                buf = torch.empty_like(slot[:, :, lo - k:lo, :]).contiguous()
                ops.append(dist.P2POp(dist.irecv, buf, up, self.group))
                recvs.append((buf, lo - k, lo))
        if down < self.world:
            k = self.planners[down].halo[0]
            if k:
                send = slot[:, :, hi - k:hi, :].contiguous()
                ops.append(dist.P2POp(dist.isend, send, down, self.group))
                nbytes += send.numel() * 4
            k = self.me.halo[1]
            if k:
                buf = torch.empty_like(slot[:, :, hi:hi + k, :]).contiguous()
                ops.append(dist.P2POp(dist.irecv, buf, down, self.group))
                recvs.append((buf, hi, hi + k))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for buf, a, b in recvs:
            slot[:, :, a:b, :].copy_(buf)
        self.halo_bytes_per_iteration = nbytes
        return nbytes

    def rollout(self, x0, series, iterations, n_out=1):
        """
        x0: full (N, C, H, W) state (every rank holds at least its band + halo rows); series: (iterations*n_out, N, C, H,
        W) on this rank, of which the band rows are valid afterwards.  The exchange of iteration t's last output runs
        before iteration t+1 starts.
        """
        for t in range(iterations):
            src = x0 if t == 0 else series[t * n_out - 1]
            outs = [series[t * n_out + k] for k in range(n_out)]
            self.forward(src, outs)
            if t + 1 < iterations:
                self.exchange(outs[-1])
        return series


def make_planners(ops, buffers, H, world):
    return [BandPlanner(ops, buffers, b) for b in band_bounds(H, world)]


def halo_summary(planner, N, C, W):
    top, bottom = planner.halo
    per_row = N * C * W * 4
    return {'halo_rows_top': int(top), 'halo_rows_bottom': int(bottom), 'bytes_per_row': per_row,
            'recv_bytes_per_iteration': int((top + bottom) * per_row)}


def bind_host_near_gpu(device_index):
    """
    Restrict this process to the CPUs NVML reports as local to CUDA device `device_index` (one process per GPU: pinned
    host buffers are then allocated on, and filled through, the socket the GPU hangs off -- on a two-socket box a D2H that
    crosses the inter-socket link loses bandwidth to every other rank doing the same).  Returns the number of CPUs kept,
    0 when nothing was changed (no NVML, no affinity support, or the mask would be empty).
    """
    import os
    try:
        import pynvml
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        local = set(64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1)
        keep = sorted(local & os.sched_getaffinity(0))
        if not keep or len(keep) == len(os.sched_getaffinity(0)):
            return 0
        os.sched_setaffinity(0, keep)
        return len(keep)
    except Exception:       # affinity is an optimisation: never a reason to fail
        return 0


def _nccl_library():
    """torch's bundled NCCL (the same shared object torch.distributed uses), loaded by the C library with dlopen."""
    import os
    try:
        import nvidia.nccl
        cand = os.path.join(list(nvidia.nccl.__path__)[0], 'lib', 'libnccl.so.2')
        if os.path.exists(cand):
            return cand
    except ImportError:
        pass
    return 'libnccl.so.2'


def create_native_comm(dist, rank, world):
    """The library's own NCCL communicator over the ranks of an initialised torch.distributed group (collective: every rank
    calls it).  torch.distributed only carries the 128-byte NCCL id from rank 0 to the others."""
    import ctypes
    import torch
    lib = nat.lib()
    path = _nccl_library().encode()
    ident = (ctypes.c_char * 128)()
    if rank == 0:
        nat.check(lib.dlwp_comm_unique_id(path, ident), 'dlwp_comm_unique_id')
    t = torch.tensor(list(bytes(ident)), dtype=torch.uint8, device='cuda')
    dist.broadcast(t, 0)
    ident = (ctypes.c_char * 128).from_buffer_copy(bytes(t.cpu().tolist()))
    comm = ctypes.c_void_p()
    nat.check(lib.dlwp_comm_create(path, rank, world, ident, ctypes.byref(comm)), 'dlwp_comm_create')
    return comm


class LatBandEngine(object):
    """
    GPU implementation: a row-windowed DlwpPlan per rank; the rollout loop INCLUDING the halo exchange runs inside the C
    library (dlwp_rollout_latband: pack rows -> ncclGroupStart / ncclSend+ncclRecv up and down / ncclGroupEnd -> unpack),
    captured as one CUDA graph.  torch.distributed is only used once, to broadcast the NCCL unique id.
    `native=False` keeps the torch.distributed driver above (what the gloo CPU tests exercise).

    halo = 'p2p': the exchange runs over peer memory instead -- the last conv's epilogue stores the neighbours' halo rows
    straight into their (IPC-mapped) input images over NVLink and two one-thread kernels count arrivals
    (dlwp_plan_halo_*); 'nccl': the grouped SendRecv; 'auto' (default): p2p when the plan qualifies and the handles can be
    exchanged, else nccl.  `self.halo` says which one is in use.
    """

    def __init__(self, model, batch, rank, world, dist=None, impl=None, native=True, halo='auto'):
        import ctypes
        from .engine import CompiledNet, Lowering
        low = Lowering(model)
        H = low.buffers[[i for i, b in enumerate(low.buffers) if b['kind'] == nat.BUF_INPUT][0]]['H']
        self.planners = make_planners(low.ops, low.buffers, H, world)
        self.me = self.planners[rank]
        self.net = CompiledNet(model, batch, impl=impl, row_windows=self.me.windows)
        if self.net.max_batch < batch:
            raise ValueError('batch %d does not fit the plan (%d)' % (batch, self.net.max_batch))
        if not self.net.can_rollout():
            raise ValueError('latitude-band rollout needs output shape == input shape')
        self.driver = LatBandRollout(H, rank, world, self.planners,
                                     lambda src, outs: self.net.forward_into(src, outs), dist=dist)
        self.rank, self.world, self.H = rank, world, H
        self._dist = dist
        self.n_out = self.net.n_outputs
        self.native = bool(native)
        self.comm = ctypes.c_void_p()
        up = self.planners[rank - 1] if rank > 0 else None
        down = self.planners[rank + 1] if rank + 1 < world else None
        self.info = nat.BandInfo(rank, world, self.me.band[0], self.me.band[1], self.me.halo[0], self.me.halo[1],
                                 up.halo[1] if up else 0, down.halo[0] if down else 0)
        self.halo = 'nccl'
        if self.native and world > 1 and halo in ('auto', 'p2p'):
            ok = self._connect_peers(dist, rank, world, strict=(halo == 'p2p'))
            if ok is None:       # the handles could not be mapped on some rank: a fresh plan, and the NCCL exchange
                self.net.close()
                self.net = CompiledNet(model, batch, impl=impl, row_windows=self.me.windows)
                ok = False
            self.halo = 'p2p' if ok else 'nccl'
        if self.native and world > 1 and self.halo == 'nccl':
            self.comm = create_native_comm(dist, rank, world)

    def _connect_peers(self, dist, rank, world, strict=False):
        """Enable the peer-memory halo on this rank's plan and map the neighbours' images (CUDA IPC handles swapped with one
        all_gather).  Every rank must reach the same verdict, so the local outcome is agreed on with an all_reduce(MIN)."""
        import ctypes
        import torch
        lib = nat.lib()
        ok = lib.dlwp_plan_halo_enable(self.net.plan) == 0
        mine = (ctypes.c_char * 192)()
        if ok:
            ok = lib.dlwp_plan_halo_export(self.net.plan, mine) == 0
        flag = torch.tensor([1 if ok else 0], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not int(flag.item()):
            if strict:
                raise RuntimeError('peer-memory halo unavailable: ' + nat.lib().dlwp_last_error_string().decode())
            return False
        t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device='cuda')
        allh = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allh, t)
        good = 1
        for which, nb in ((0, rank - 1), (1, rank + 1)):
            if 0 <= nb < world:
                buf = (ctypes.c_char * 192).from_buffer_copy(bytes(allh[nb].cpu().tolist()))
                if lib.dlwp_plan_halo_import(self.net.plan, which, buf) != 0:
                    good = 0
        flag = torch.tensor([good], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not int(flag.item()):
            if strict:
                raise RuntimeError('peer-memory halo: cudaIpcOpenMemHandle failed on some rank: ' +
                                   nat.lib().dlwp_last_error_string().decode())
            return None
        return True

    def close(self, destroy_comm=False):
        """ncclCommDestroy is collective-like (it can block until every rank calls it); by default the communicator is
        left to process teardown."""
        if self.comm and destroy_comm:
            nat.lib().dlwp_comm_destroy(self.comm)
        self.comm = None
        self.net.close()

    def rollout_device(self, x0, iterations, out=None, use_graph=True):
        """
        x0: CUDA (N, C, H, W), identical on every rank (or at least valid on band + halo rows).  Returns the full-shape
        series tensor of this rank; only rows `band` are meaningful.
        """
        import ctypes
        import torch
        shape = (iterations * self.n_out,) + tuple(x0.shape)
        series = out if out is not None else torch.empty(shape, dtype=torch.float32, device=x0.device)
        if not self.native:
            return self.driver.rollout(x0, series, iterations, self.n_out)
        self.net.sync_weights()
        nat.check(nat.lib().dlwp_rollout_latband(self.net.plan, self.comm, x0.shape[0], x0.data_ptr(), series.data_ptr(),
                                                 int(iterations), ctypes.byref(self.info), 1 if use_graph else 0,
                                                 self.net._stream()), 'dlwp_rollout_latband')
        return series

    def rollout_host(self, x0, iterations, out=None, d2h_group=0):
        """
        numpy in -> numpy out through dlwp_rollout_latband_host: x0 (N, C, H, W), identical on every rank; returns this
        rank's band of every state, (iterations * n_out, N, C, band_rows, W), in pinned host memory (`out`: a pinned torch
        tensor of that shape to reuse).  The band rows of finished steps travel D2H while the next steps compute.
        """
        import ctypes
        import torch
        if not self.native:
            raise RuntimeError('rollout_host needs the native lat-band rollout')
        x0 = np.ascontiguousarray(x0, np.float32)
        lo, hi = self.me.band
        shape = (int(iterations) * self.n_out,) + tuple(x0.shape[:2]) + (hi - lo, x0.shape[3])
        buf = out if out is not None else torch.empty(shape, dtype=torch.float32, pin_memory=True)
        if tuple(buf.shape) != shape or buf.dtype != torch.float32 or not buf.is_contiguous():
            raise ValueError('out must be a contiguous float32 tensor of shape %r' % (shape,))
        self.net.sync_weights()
        if self.world > 1 and not self.comm:
            # (peer-memory halo: the library's communicator was not needed so far -- here it carries the 4-byte max|x0|
            # reduction that lets every rank upload only its own rows of x0; collective, every rank gets here together)
            self.comm = create_native_comm(self._dist, self.rank, self.world)
        nat.check(nat.lib().dlwp_rollout_latband_host(self.net.plan, self.comm, x0.shape[0], x0.ctypes.data,
                                                      buf.data_ptr(), int(iterations), ctypes.byref(self.info),
                                                      int(d2h_group)), 'dlwp_rollout_latband_host')
        return buf.numpy()

    def uploaded_rows(self):
        """Rows of x0 that `rollout_host` copies to this rank's device: the band plus the halo rows it reads."""
        lo, hi = self.me.band
        if self.world == 1:
            return 0, self.H
        return max(0, lo - self.me.halo[0]), min(self.H, hi + self.me.halo[1])

    def predict_timeseries(self, predictors, time_steps, gather=True):
        """The lat-band counterpart of `DLWPNeuralNet.predict_timeseries` (models.py:247-301) for time_dim == 1 models:
        every rank passes the SAME numpy predictors (N, C, H, W); the state is rolled forward `time_steps` times with the
        latitude split over the ranks.  gather=True: rank 0 returns the complete float32 series (steps, N, C, H, W)
        assembled from all bands (one collective gather of the per-rank bands), the other ranks return None;
        gather=False: every rank returns (its band of the series, (row_lo, row_hi))."""
        import torch
        time_steps = int(time_steps)
        if time_steps < 1:
            raise ValueError("time_steps must be an int > 0")
        lo, hi = self.me.band
        if not gather or self.world == 1:   # host in, host out: the pipelined native path
            band = self.rollout_host(predictors, time_steps)
            return (band, (lo, hi)) if not gather else band
        x0 = torch.from_numpy(np.ascontiguousarray(predictors, np.float32)).cuda()
        series = self.rollout_device(x0, time_steps)
        band = series[:, :, :, lo:hi, :].contiguous()
        import torch.distributed as dist
        rows = [p.band[1] - p.band[0] for p in self.planners]
        pad = max(rows)                                  # dist.gather wants equal shapes: pad the shorter bands
        buf = torch.zeros(band.shape[:3] + (pad, band.shape[4]), dtype=torch.float32, device='cuda')
        buf[:, :, :, :hi - lo] = band
        parts = [torch.empty_like(buf) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(buf, parts, dst=0)
        if self.rank != 0:
            return None
        return torch.cat([p[:, :, :, :r] for p, r in zip(parts, rows)], dim=3).cpu().numpy()

    def band_to_host(self, series):
        """This rank's band of the series as a pinned numpy array (steps, N, C, band_rows, W)."""
        import torch
        lo, hi = self.me.band
        band = series[:, :, :, lo:hi, :].contiguous()
        host = torch.empty(band.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(band, non_blocking=True)
        torch.cuda.synchronize()
        return host.numpy()
