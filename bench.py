#!/usr/bin/env python
"""
bench.py -- forecast-steps/sec of DLWP's predict_timeseries rollout (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Workload (BASELINE.json configs[1]): "Net A" -- PeriodicPadding2D((0,2)) + ZeroPadding2D((2,0)) + Conv2D(32,3,d=2,tanh)
+ PeriodicPadding2D + ZeroPadding2D + Conv2D(6,5,linear) on a (B, 6, 91, 180) fp32 state, K feedback steps; synthetic
inputs x0 = RandomState(0).randn, weights = Keras glorot_uniform restated with RandomState(1), zero biases (BASELINE.md
section 3).  One "step" = one rollout step over the whole batch; value = B * K / time in forecast-steps/s.

* `value`      device-resident rollout (x0 and the series in HBM), one CUDA-graph launch of K steps, CUDA events.
* `e2e`        the reference-facing call DLWPNeuralNet.predict_timeseries(numpy) -> numpy: H2D of x0, rollout, D2H of
               every step's state, wall clock around the call.  N > 1 (latitude bands): LatBandEngine.rollout_host --
               every rank uploads the rows of x0 it reads and receives its band of every state (dlwp_rollout_latband_host).
* N > 1        the north-star partition: latitude bands, halo rows over NVLink peer memory (--halo nccl: grouped SendRecv),
               --scaling weak (default: B forecasts per GPU) | strong; --parallel batch = independent forecasts per GPU.
* --workload net_b [--precision bf16]: the skip U-Net on 12x180x360 (BASELINE.json configs[2], configs[3] with --gpus N).
* `roofline`   the dominant kernel (conv 32->6 5x5) timed alone with CUDA events: SURVEY.md 8(d) algorithmic bytes / time
               against the measured HBM copy peak (MEASURED_PEAKS.json).
* `cpu_baseline` / `--impl reference`: the reference's rollout loop restated in oracle/ around a torch-CPU fp32 forward
               (Keras/TensorFlow are not installable offline), all host threads, on a bounded sample.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

STATE = (6, 91, 180)
BYTES_PER_SAMPLE_STEP = 5005784      # sum over conv layers of 4*Ho*Wo*(Cin+Cout) + 4*(weights+bias)
FLOP_PER_SAMPLE_STEP = 213857280


def conv_alg_bytes(n, cin, cout, k, ho=91, wo=180):
    return 4 * n * ho * wo * (cin + cout) + 4 * (k * k * cin * cout + cout)


def measured_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


def ncu_traffic(kernel, batch):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    `ncu --set full` capture (profiles/r02_ncu_traffic.json, taken at the bench batch); None if there is no capture for
    this kernel / batch."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r02_ncu_traffic.json')
    try:
        table = json.load(open(path))
    except (IOError, OSError, ValueError):
        return None
    for row in table:
        if row.get('batch') == batch and kernel.startswith(row.get('kernel_prefix', '\0')) and row.get('layer', '') in kernel:
            return row.get('dram_bytes_per_launch')
    return None


def make_inputs(batch):
    return np.random.RandomState(0).standard_normal((batch,) + STATE).astype(np.float32)


def net_a_layers():
    """First and last conv blocks of examples/train.py:159-169, 211-219 as DLWPNeuralNet.build_model layer tuples."""
    cf = 'channels_first'
    return (
        ('PeriodicPadding2D', ((0, 2),), {'data_format': cf, 'input_shape': STATE}),
        ('ZeroPadding2D', ((2, 0),), {'data_format': cf}),
        ('Conv2D', (32, 3), {'dilation_rate': 2, 'padding': 'valid', 'activation': 'tanh', 'data_format': cf}),
        ('PeriodicPadding2D', ((0, 2),), {'data_format': cf}),
        ('ZeroPadding2D', ((2, 0),), {'data_format': cf}),
        ('Conv2D', (STATE[0], 5), {'padding': 'valid', 'activation': 'linear', 'data_format': cf}),
    )


def net_a_weights():
    """Keras default glorot_uniform restated with RandomState(1), (kh,kw,Cin,Cout) order per layer; zero biases."""
    rng = np.random.RandomState(1)
    ws = []
    for kh, kw, cin, cout in ((3, 3, STATE[0], 32), (5, 5, 32, STATE[0])):
        limit = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
        ws += [rng.uniform(-limit, limit, size=(kh, kw, cin, cout)).astype(np.float32), np.zeros(cout, np.float32)]
    return ws


def oracle_net():
    """cpu_baseline / --impl reference only: the oracle's Net A with the same weights."""
    from oracle import layers as OL
    net = OL.OSequential(net_a_layers())
    net.set_weights(net_a_weights())
    return net


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  The sampler
    is started ahead of the region (nvidia-smi needs a moment to come up); `mark()` brackets the region and only samples
    taken between the marks -- while the kernels run -- enter the summary."""
    FIELDS = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index, period_ms=20):
        self.rows = []
        self.proc = None
        self.index = index
        self.period_ms = period_ms
        self.marks = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:      # first sample in hand before the region starts
                time.sleep(0.01)
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def mark(self):
        self.marks.append(time.time())

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(2.5 * self.period_ms / 1000.0)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        lo, hi = (self.marks[0], self.marks[-1] + self.period_ms / 1000.0) if len(self.marks) >= 2 else (0, float('inf'))
        inside = [r for t, r in self.rows if lo <= t <= hi] or [r for _, r in self.rows[-2:]]
        for r in inside:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm), 'period_ms': self.period_ms}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference(n_samples, steps, warmup):
    """The reference's rollout loop (oracle.rollout) around a torch-CPU forward; returns (forecast-steps/s, cores)."""
    from oracle import rollout as OR
    from oracle import torch_cpu as OT
    import torch
    model = OT.KerasLikeModel(oracle_net())
    x0 = make_inputs(n_samples)
    cores = OT.use_all_cores()
    # oneDNN does not always scale to every hardware thread on these small convolutions: give the baseline the thread
    # count that serves it best (probe one step at all / half / quarter of the cores, keep the fastest)
    best = None
    for nt in sorted({cores, max(1, cores // 2), max(1, cores // 4), min(cores, 32)}, reverse=True):
        torch.set_num_threads(nt)
        OR.neuralnet_predict_timeseries(model.predict, x0, 1)
        t0 = time.perf_counter()
        OR.neuralnet_predict_timeseries(model.predict, x0, 1)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, nt)
    cores = best[1]
    torch.set_num_threads(cores)
    if warmup > 0:
        OR.neuralnet_predict_timeseries(model.predict, x0, warmup)
    t0 = time.perf_counter()
    OR.neuralnet_predict_timeseries(model.predict, x0, steps)
    dt = time.perf_counter() - t0
    return n_samples * steps / dt, cores, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    if args.workload == 'net_b':
        _, onet = build_net_b_oracle_only()
        v, cores, _ = net_b_cpu_reference(onet, 4, 1, 1)
        n = int(max(1, min(16, 120.0 * v / max(1, args.steps + args.warmup))))
        value, cores, dt = net_b_cpu_reference(onet, n, args.steps, args.warmup)
        sample = '%d of %d samples x %d steps (Keras-style batch_size=32 chunks)' % (n, args.batch, args.steps)
        print(json.dumps({
            'impl': 'reference', 'metric': 'forecast_steps_per_sec', 'value': value, 'unit': 'forecast-steps/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'net_b_skip_unet_12x180x360_rollout', 'batch_per_gpu': args.batch, 'timed_sample_batch': n,
                       'note': 'reference rollout loop (DLWP/model/models.py:414-452 restated in oracle/) + torch-CPU fp32 '
                               'forward; Keras/TensorFlow are not installable offline'},
            'cpu_baseline': {'value': value, 'unit': 'forecast-steps/s', 'cores': cores, 'kind': 'port', 'host_cores': os.cpu_count(), 'sample': sample},
            'e2e': {'value': value, 'unit': 'forecast-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}))
        return
    # bounded sample: probe one step on 8 samples, then size the sample so the whole run stays within ~2 minutes
    v, cores, _ = cpu_reference(8, 1, 1)
    budget_s = 120.0
    n = int(max(1, min(32, budget_s * v / max(1, args.steps + args.warmup))))
    value, cores, dt = cpu_reference(n, args.steps, args.warmup)
    sample = '%d of %d samples x %d steps (Keras-style batch_size=32 chunks)' % (n, args.batch, args.steps)
    line = {
        'impl': 'reference', 'metric': 'forecast_steps_per_sec', 'value': value, 'unit': 'forecast-steps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'net_a_6x91x180_rollout', 'batch_per_gpu': args.batch, 'timed_sample_batch': n,
                   'note': 'reference rollout loop (DLWP/model/models.py:247-301 restated in oracle/) + torch-CPU fp32 '
                           'forward; Keras/TensorFlow are not installable offline'},
        'cpu_baseline': {'value': value, 'unit': 'forecast-steps/s', 'cores': cores, 'kind': 'port', 'host_cores': os.cpu_count(), 'sample': sample},
        'e2e': {'value': value, 'unit': 'forecast-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def build_model():
    from dlwp_b200.model import DLWPNeuralNet
    dlwp = DLWPNeuralNet(is_convolutional=True, is_recurrent=False, time_dim=1, scaler_type=None, scale_targets=False)
    dlwp.build_model(net_a_layers(), loss='mse', optimizer='adam')
    dlwp.model.set_weights(net_a_weights())
    return dlwp


def time_single_kernel(torch, nat, batch, cin, cout, k, d, act, iters):
    """Average duration (ms) of one conv layer launched alone, CUDA events on the launching stream."""
    import ctypes
    rng = np.random.RandomState(3)
    x = torch.from_numpy(rng.standard_normal((batch, cin, 91, 180)).astype(np.float32)).cuda()
    w = torch.from_numpy((0.05 * rng.standard_normal((k, k, cin, cout))).astype(np.float32)).cuda()
    b = torch.zeros(cout, device='cuda')
    y = torch.empty((batch, cout, 91, 180), device='cuda')
    desc = nat.ConvDesc(N=batch, Cin=cin, H=91, W=180, Cout=cout, kh=k, kw=k, dil_h=d, dil_w=d, pad_t=2, pad_b=2,
                        pad_l=2, pad_r=2, pad_mode_h=0, pad_mode_w=1, act=act, pre_op=0, rowwise=0, impl=0, reserved=0, row_begin=0, row_end=0,
                        x_stride_n=cin * 91 * 180, x_stride_c=91 * 180, x_stride_h=180, y_stride_n=cout * 91 * 180,
                        y_stride_c=91 * 180, y_stride_h=180)
    lib = nat.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    name = lib.dlwp_conv2d_impl_name(ctypes.byref(desc)).decode()
    for _ in range(3):
        nat.check(lib.dlwp_conv2d_fwd(ctypes.byref(desc), x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(),
                                      stream))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.dlwp_conv2d_fwd(ctypes.byref(desc), x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, name


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from dlwp_b200 import _native as nat
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    lib = nat.lib()
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    dlwp = build_model()
    eng = dlwp.model.engine(B)
    assert eng.max_batch >= B, 'batch does not fit the plan'
    # every rank rolls its own B samples forward (independent forecasts: no data-path collective)
    x0 = np.random.RandomState(rank).standard_normal((B,) + STATE).astype(np.float32)
    xd = torch.from_numpy(x0).cuda()
    series = torch.empty((K, B) + STATE, dtype=torch.float32, device='cuda')
    warm = torch.empty((W, B) + STATE, dtype=torch.float32, device='cuda')

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.rollout_device(xd, W, use_graph=False, out=warm)          # W untimed warm-up steps
    eng.rollout_device(xd, K, use_graph=True, out=series)         # graph capture + first replay (untimed)
    barrier()
    launches0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        clocks.mark()
        e0.record()
        eng.rollout_device(xd, K, use_graph=True, out=series)     # EXACTLY K timed steps
        e1.record()
        barrier()
        clocks.mark()
    ms = e0.elapsed_time(e1)
    launches = nat.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * K / (ms * 1e-3)

    # ---- end to end through the reference-facing API: numpy in -> numpy out, copies inside the timed region ----------
    Ke = min(K, args.e2e_steps)
    x0_pinned = torch.from_numpy(x0).pin_memory().numpy()
    y = dlwp.predict_timeseries(x0_pinned, Ke)                    # warm-up: sizes the device series and the pinned
    del y                                                         # host pool exactly as the timed call needs them
    barrier()
    t0 = time.perf_counter()
    y = dlwp.predict_timeseries(x0_pinned, Ke)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert y.shape == (Ke, B) + STATE
    if world > 1:
        t = torch.tensor([dt], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_value = world * B * Ke / dt
    checksum = float(np.abs(y[-1]).mean())
    slot_bytes = B * int(np.prod(STATE)) * 4
    del y

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: each layer of the plan launched alone, CUDA events (rank 0) ----------------------
    peaks, peak_kind = measured_peaks()
    it = max(10, min(50, K))
    tc = eng.uses_tensor_cores()
    layers = [('conv1 6->32 3x3 d2 tanh', 6, 32, 3), ('conv2 32->6 5x5 linear', 32, 6, 5)]
    timed = []
    for i, (name, cin, cout, k) in enumerate(layers):
        ms_i = eng.profile_op(B, i, it)
        alg = conv_alg_bytes(B, cin, cout, k)
        tc_name = 'conv_sw_kernel '
        timed.append({'kernel': (tc_name if tc else 'conv_ffma_kernel ') + name, 'ms_per_launch': ms_i,
                      'algorithmic_bytes_per_launch': alg, 'achieved_gbs': alg / (ms_i * 1e-3) / 1e9,
                      'useful_tflops': 2.0 * B * 91 * 180 * cin * cout * k * k / (ms_i * 1e-3) / 1e12})
    dom = max(timed, key=lambda r: r['ms_per_launch'])
    step_gbs = BYTES_PER_SAMPLE_STEP * B / (ms / K * 1e-3) / 1e9
    roofline = {
        'kernel': dom['kernel'], 'bound': 'hbm', 'achieved': dom['achieved_gbs'], 'peak': peaks['hbm_gbs'],
        'unit': 'GB/s', 'frac': dom['achieved_gbs'] / peaks['hbm_gbs'], 'traffic': ncu_traffic(dom['kernel'], B),
        'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peak_kind == 'measured' else 'fallback 6650 GB/s',
        'algorithmic_bytes_per_launch': dom['algorithmic_bytes_per_launch'], 'ms_per_launch': dom['ms_per_launch'],
        'math': 'tcgen05 f16 hi/lo split x3 -> fp32 TMEM accumulators' if tc else 'fp32 FFMA2',
        'layers': timed,
        'step': {'algorithmic_bytes': BYTES_PER_SAMPLE_STEP * B, 'ms': ms / K, 'achieved_gbs': step_gbs,
                 'frac': step_gbs / peaks['hbm_gbs'],
                 'useful_tflops': FLOP_PER_SAMPLE_STEP * B / (ms / K * 1e-3) / 1e12},
    }

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        v, cores, dtc = cpu_reference(32, 4, 1)
        cpu = {'value': v, 'unit': 'forecast-steps/s', 'cores': cores, 'kind': 'port', 'host_cores': os.cpu_count(),
               'sample': '32 of %d samples x 4 steps (1 warm-up step), torch-CPU fp32 forward inside the reference '
                         'rollout loop' % B}

    line = {
        'metric': 'forecast_steps_per_sec', 'value': value, 'unit': 'forecast-steps/s', 'n_gpus': world, 'steps': K,
        'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if (args.precision == 'bf16' and tc) else 'f32', 'data': 'synthetic',
        'config': {'workload': 'net_a_6x91x180_rollout (BASELINE.json configs[1])', 'batch_per_gpu': B,
                   'precision': args.precision,
                   'global_batch': B * world, 'state': list(STATE), 'math': args.math, 'parallelism': 'independent forecasts per GPU'
                   if world > 1 else 'single GPU',
                   'l2': 'per-step working set %.0f MB (state in + 32-ch activation + state out) > 126 MB L2; no flush'
                         % ((6 + 32 + 6) * 91 * 180 * 4 * B / 1e6),
                   'timed_region': 'one CUDA-graph replay of %d steps, inputs resident in HBM' % K,
                   'e2e_steps': Ke, 'checksum_mean_abs_last_state': checksum},
        'e2e': {'value': e2e_value, 'unit': 'forecast-steps/s', 'h2d_bytes_per_step': slot_bytes / Ke,
                'd2h_bytes_per_step': slot_bytes, 'api': 'DLWPNeuralNet.predict_timeseries(numpy)->numpy',
                'steps': Ke, 'seconds': dt},
        'gpu_launches': launches,
        'clocks': clocks.summary(),
        'roofline': roofline,
        'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


HALO_TEXT = {
    'p2p': 'no collective on the data path: the last conv\'s epilogue stores the rows the neighbours need straight into '
           'their input images over NVLink (CUDA IPC peer memory); two one-thread kernels per step count arrivals; all '
           'captured in one CUDA graph',
    'nccl': 'one grouped NCCL SendRecv per step (ncclSend/ncclRecv inside the C library) between one packing and one '
            'unpacking kernel, captured with the band kernels in one CUDA graph'}


def run_latband(args, rank, world, local_rank):
    """
    N > 1: the north-star partition -- every GPU owns a latitude band of ALL forecasts in flight and exchanges a 4-row halo
    of the state with its neighbours once per step (one grouped NCCL SendRecv).  --scaling weak (default): the global batch
    grows with the GPU count (B per GPU: B * N forecasts, each GPU computes its band of all of them -- per-GPU work fixed up
    to the halo rows it recomputes); --scaling strong: the global batch stays B.
    """
    import torch
    import torch.distributed as dist
    from dlwp_b200 import _native as nat
    from dlwp_b200.parallel import LatBandEngine, bind_host_near_gpu, halo_summary
    torch.cuda.set_device(local_rank)
    bound_cpus = bind_host_near_gpu(local_rank)           # pinned host buffers on the GPU's own socket
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    K, W = args.steps, max(args.warmup, 3)
    B = args.batch * world if args.scaling == 'weak' else args.batch
    free, _ = torch.cuda.mem_get_info()
    cap = int(0.5 * free / ((K + 2) * int(np.prod(STATE)) * 4))   # every rank holds full-shape series slots
    capped = B > cap
    B = max(world, min(B, cap))
    dlwp = build_model()
    eng = LatBandEngine(dlwp.model, B, rank, world, dist=dist, halo=args.halo)
    x0 = make_inputs(B)                                    # the same global state on every rank
    xd = torch.from_numpy(x0).cuda()
    series = torch.empty((K, B) + STATE, dtype=torch.float32, device='cuda')

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    eng.rollout_device(xd, W, out=series[:W], use_graph=False)   # W untimed warm-up steps (also warms NCCL P2P channels)
    eng.rollout_device(xd, K, out=series)                  # graph capture + first replay (untimed)
    barrier()
    launches0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        clocks.mark()
        e0.record()
        eng.rollout_device(xd, K, out=series)              # EXACTLY K timed steps, K-1 halo exchanges
        e1.record()
        barrier()
        clocks.mark()
    t = torch.tensor([e0.elapsed_time(e1)], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = nat.launch_count() - launches0
    value = B * K / (ms * 1e-3)

    # self-check of the exchange (local, no extra collective): on a small batch this rank's band of a graph-replayed
    # lat-band rollout must equal the same rows of the single-domain rollout computed on this GPU, bit for bit
    vb, vk = 2, 5
    veng = eng                                            # same plan and communicator, smaller batch
    vx = xd[:vb].contiguous()
    vs = torch.full((vk, vb) + STATE, float('nan'), device='cuda')
    veng.rollout_device(vx, vk, out=vs)
    vref = dlwp.model.engine(vb).rollout_device(vx, vk, use_graph=False)
    torch.cuda.synchronize()
    lo_, hi_ = veng.me.band
    vhost = veng.rollout_host(vx.cpu().numpy(), vk)        # the host-buffer path must deliver the same band
    same = torch.equal(vs[:, :, :, lo_:hi_], vref[:, :, :, lo_:hi_]) and \
        np.array_equal(vhost, vref[:, :, :, lo_:hi_].cpu().numpy())
    vflag = torch.tensor([1 if same else 0], device='cuda')
    dist.all_reduce(vflag, op=dist.ReduceOp.MIN)
    verified = bool(int(vflag.item()))

    # end to end: H2D of x0, rollout, D2H of this rank's band of every state (host concatenation along H is free)
    Ke = min(K, args.e2e_steps)
    x0_pinned = torch.from_numpy(x0).pin_memory().numpy()
    del series                                            # the host path keeps its own device series inside the plan
    torch.cuda.empty_cache()
    band = eng.rollout_host(x0_pinned, Ke)                # warm-up: sizes the device series and the pinned host pool
    del band
    barrier()
    t0 = time.perf_counter()
    band = eng.rollout_host(x0_pinned, Ke)                # numpy in -> this rank's band of every state, numpy out
    dt = time.perf_counter() - t0
    assert band.shape == (Ke, B, STATE[0], eng.me.band[1] - eng.me.band[0], STATE[2]) and np.isfinite(band[-1]).all()
    t = torch.tensor([dt], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    rows = eng.me.band[1] - eng.me.band[0]
    up_rows = eng.uploaded_rows()                         # per rank: the band plus the halo rows it reads
    if rank == 0:
        halo = halo_summary(eng.planners[min(1, world - 1)], B, STATE[0], STATE[2])
        per_dir = halo['bytes_per_row'] * 4
        line = {
            'metric': 'forecast_steps_per_sec', 'value': value, 'unit': 'forecast-steps/s', 'n_gpus': world, 'steps': K,
            'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'net_a_6x91x180_rollout (BASELINE.json configs[1])', 'global_batch': B,
                       'batch_per_gpu': B / world, 'global_batch_capped_by_memory': capped,
                       'state': list(STATE), 'parallelism': 'latband%d (91 latitude rows split over %d GPUs, halo 4 rows)'
                       % (world, world), 'bands': [list(p.band) for p in eng.planners],
                       'bands_equal_single_domain_bitwise': verified, 'host_cpus_bound_near_gpu': bound_cpus,
                       'halo': {'rows_per_side': 4, 'bytes_per_neighbour_per_direction_per_step': per_dir,
                                'exchange': eng.halo,
                                'collective': HALO_TEXT[eng.halo],
                                'link_time_us_at_770GBs': per_dir / 770e9 * 1e6,
                                'fraction_of_step_time': per_dir / 770e9 / (ms / K * 1e-3)},
                       'l2': 'per-step working set >> L2 at this batch; no flush', 'e2e_steps': Ke},
            'e2e': {'value': B * Ke / dt, 'unit': 'forecast-steps/s', 'h2d_bytes_per_step': B * STATE[0] * (up_rows[1] - up_rows[0]) * STATE[2] * 4 / Ke,
                    'd2h_bytes_per_step': B * STATE[0] * rows * STATE[2] * 4,
                    'api': 'LatBandEngine.rollout_host(numpy) -> numpy (per-rank band; strided D2H pipelined behind the steps)', 'steps': Ke, 'seconds': dt},
            'gpu_launches': launches, 'clocks': clocks.summary(),
            'roofline': None, 'cpu_baseline': None,
        }
        print(json.dumps(line))
    dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# Net B: the skip U-Net of examples/train_functional.py:248-275 on the 12-channel 1-degree grid (BASELINE.json configs[2],
# lat-band over N GPUs: configs[3]).  python bench.py --workload net_b [--precision bf16] [--batch 16]
# ---------------------------------------------------------------------------------------------------------------------
NET_B_STATE = (12, 180, 360)
NET_B_BYTES = {'fp32': 62111600.0, 'bf16': 31266800.0}     # SURVEY.md 8(d), per sample-step
NET_B_FLOP = 4678041600.0


def build_net_b():
    from tests.helpers import build_functional_pair
    dlwp, onet = build_functional_pair(NET_B_STATE, skip=True, integration_steps=1, seed=1, bias_scale=0.02)
    return dlwp, onet


def build_net_b_oracle_only():
    from oracle import layers as OL
    onet = OL.OFunctionalNet(NET_B_STATE, skip_connections=True, integration_steps=1)
    OL.init_weights(onet.conv_layers, seed=1, bias_scale=0.02)
    return None, onet


def net_b_cpu_reference(onet, n_samples, steps, warmup):
    from oracle import rollout as OR
    from oracle import torch_cpu as OT
    model = OT.KerasLikeModel(onet)
    cores = OT.use_all_cores()
    x0 = np.random.RandomState(0).standard_normal((n_samples,) + NET_B_STATE).astype(np.float32)
    if warmup:
        OR.functional_predict_timeseries(model.predict, x0, warmup)
    t0 = time.perf_counter()
    OR.functional_predict_timeseries(model.predict, x0, steps)
    dt = time.perf_counter() - t0
    return n_samples * steps / dt, cores, dt


def run_net_b(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from dlwp_b200 import _native as nat
    torch.cuda.set_device(local_rank)
    if world > 1:
        from dlwp_b200.parallel import bind_host_near_gpu
        bind_host_near_gpu(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    K, W = args.steps, max(args.warmup, 3)
    prec = args.precision
    dlwp, onet = build_net_b()
    latband = world > 1 and args.parallel == 'latband'
    B = args.batch * world if (latband and args.scaling == 'weak') else args.batch
    x0 = np.random.RandomState(0 if latband else rank).standard_normal((B,) + NET_B_STATE).astype(np.float32)
    xd = torch.from_numpy(x0).cuda()
    series = torch.empty((K, B) + NET_B_STATE, dtype=torch.float32, device='cuda')

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if latband:
        from dlwp_b200.parallel import LatBandEngine, halo_summary
        eng = LatBandEngine(dlwp.model, B, rank, world, dist=dist, halo=args.halo)
        net = eng.net
        roll = lambda k, out, graph=True: eng.rollout_device(xd, k, out=out, use_graph=graph)
    else:
        eng = None
        net = dlwp.model.engine(B)
        assert net.max_batch >= B, 'batch does not fit the plan'
        roll = lambda k, out, graph=True: net.rollout_device(xd, k, use_graph=graph, out=out)
    roll(W, series[:W], False)
    roll(K, series)
    barrier()
    launches0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        clocks.mark()
        e0.record()
        roll(K, series)
        e1.record()
        barrier()
        clocks.mark()
    ms = e0.elapsed_time(e1)
    launches = nat.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    n_forecasts = B if latband else B * world
    value = n_forecasts * K / (ms * 1e-3)
    verified = None
    if latband:
        vb, vk = 2, 3
        vx = xd[:vb].contiguous()
        vs = torch.full((vk, vb) + NET_B_STATE, float('nan'), device='cuda')
        eng.rollout_device(vx, vk, out=vs)
        vref = dlwp.model.engine(vb).rollout_device(vx, vk, use_graph=False)
        torch.cuda.synchronize()
        lo_, hi_ = eng.me.band
        vflag = torch.tensor([1 if torch.equal(vs[:, :, :, lo_:hi_], vref[:, :, :, lo_:hi_]) else 0], device='cuda')
        dist.all_reduce(vflag, op=dist.ReduceOp.MIN)
        verified = bool(int(vflag.item()))
    # end to end
    Ke = min(K, args.e2e_steps)
    if latband:
        x0_pinned = torch.from_numpy(x0).pin_memory().numpy()
        del series
        torch.cuda.empty_cache()
        band = eng.rollout_host(x0_pinned, Ke)
        del band
        barrier()
        t0 = time.perf_counter()
        band = eng.rollout_host(x0_pinned, Ke)
        dt = time.perf_counter() - t0
        rows = eng.me.band[1] - eng.me.band[0]
        assert band.shape == (Ke, B, NET_B_STATE[0], rows, NET_B_STATE[2]) and np.isfinite(band[-1]).all()
        del band
        d2h = B * NET_B_STATE[0] * rows * NET_B_STATE[2] * 4
        h2d = B * NET_B_STATE[0] * (eng.uploaded_rows()[1] - eng.uploaded_rows()[0]) * NET_B_STATE[2] * 4
        api = 'LatBandEngine.rollout_host(numpy) -> numpy (per-rank band; strided D2H pipelined behind the steps)'
    else:
        x0_pinned = torch.from_numpy(x0).pin_memory().numpy()
        dlwp.predict_timeseries(x0_pinned, Ke)
        barrier()
        t0 = time.perf_counter()
        y = dlwp.predict_timeseries(x0_pinned, Ke)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert y.shape == (Ke, B) + NET_B_STATE and np.isfinite(y[-1]).all()
        del y
        d2h = h2d = B * int(np.prod(NET_B_STATE)) * 4
        api = 'DLWPFunctional.predict_timeseries(numpy)->numpy'
    if world > 1:
        t = torch.tensor([dt], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_kind = measured_peaks()
    tc = net.uses_tensor_cores()
    s_act = 2 if (prec == 'bf16' and tc) else 4
    timed = []
    for i, op in enumerate(net.low.ops if not latband else [o for o, w in zip(net.low.ops, net.row_windows) if w is not None]):
        if op['kind'] != nat.OP_CONV or latband:
            continue
        ms_i = net.profile_op(B, i, 5)
        ob = net.low.buffers[op['dst']]
        alg = s_act * B * ob['H'] * ob['W'] * (op['src_c'] + op['Cout']) + 4 * (op['kh'] * op['kw'] * op['src_c'] * op['Cout'] + op['Cout'])
        timed.append({'kernel': 'conv_sw_kernel %d->%d %dx%d @%dx%d' % (op['src_c'], op['Cout'], op['kh'], op['kw'], ob['H'], ob['W']),
                      'ms_per_launch': ms_i, 'algorithmic_bytes_per_launch': alg, 'achieved_gbs': alg / (ms_i * 1e-3) / 1e9,
                      'useful_tflops': 2.0 * B * ob['H'] * ob['W'] * op['src_c'] * op['Cout'] * op['kh'] * op['kw'] / (ms_i * 1e-3) / 1e12})
    step_bytes = NET_B_BYTES['bf16' if s_act == 2 else 'fp32'] * B
    step_gbs = step_bytes / (ms / K * 1e-3) / 1e9 / (world if latband else 1)
    roofline = None
    if timed:
        dom = max(timed, key=lambda r: r['ms_per_launch'])
        roofline = {'kernel': dom['kernel'], 'bound': 'hbm', 'achieved': dom['achieved_gbs'], 'peak': peaks['hbm_gbs'],
                    'unit': 'GB/s', 'frac': dom['achieved_gbs'] / peaks['hbm_gbs'], 'traffic': None,
                    'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peak_kind == 'measured' else 'fallback 6650 GB/s',
                    'algorithmic_bytes_per_launch': dom['algorithmic_bytes_per_launch'], 'ms_per_launch': dom['ms_per_launch'],
                    'math': ('tcgen05 bf16 x bf16 -> fp32 TMEM accumulators, one pass' if s_act == 2 else
                             'tcgen05 f16 hi/lo split x3 -> fp32 TMEM accumulators') if tc else 'fp32 FFMA2',
                    'layers': timed,
                    'step': {'algorithmic_bytes': step_bytes, 'ms': ms / K, 'achieved_gbs': step_gbs,
                             'frac': step_gbs / peaks['hbm_gbs'], 'useful_tflops': NET_B_FLOP * B / (ms / K * 1e-3) / 1e12}}
    cpu = None
    if world == 1 and not args.no_cpu:
        v, cores, _ = net_b_cpu_reference(onet, 8, 2, 1)
        cpu = {'value': v, 'unit': 'forecast-steps/s', 'cores': cores, 'kind': 'port', 'host_cores': os.cpu_count(),
               'sample': '8 of %d samples x 2 steps (1 warm-up step), torch-CPU fp32 forward inside the reference rollout loop' % B}
    config = {'workload': 'net_b_skip_unet_12x180x360_rollout (BASELINE.json configs[%d])' % (3 if latband else 2),
              'batch_per_gpu': B / world if latband else B, 'global_batch': n_forecasts, 'state': list(NET_B_STATE),
              'precision': prec, 'tensor_cores': bool(tc),
              'parallelism': ('latband%d' % world) if latband else ('independent forecasts per GPU' if world > 1 else 'single GPU'),
              'l2': 'per-step working set >> 126 MB L2 at this batch; no flush',
              'timed_region': 'one CUDA-graph replay of %d steps, inputs resident in HBM' % K, 'e2e_steps': Ke}
    if latband:
        hs = halo_summary(eng.planners[min(1, world - 1)], B, NET_B_STATE[0], NET_B_STATE[2])
        config.update({'bands': [list(p.band) for p in eng.planners], 'bands_equal_single_domain_bitwise': verified,
                       'halo': {'rows_top_bottom': [hs['halo_rows_top'], hs['halo_rows_bottom']],
                                'recv_bytes_per_step': hs['recv_bytes_per_iteration'],
                                'link_time_us_at_770GBs': hs['recv_bytes_per_iteration'] / 2 / 770e9 * 1e6,
                                'exchange': eng.halo, 'collective': HALO_TEXT[eng.halo]}})
    line = {'metric': 'forecast_steps_per_sec', 'value': value, 'unit': 'forecast-steps/s', 'n_gpus': world, 'steps': K,
            'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True,
            'scaling': args.scaling if latband else 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if s_act == 2 else 'f32', 'data': 'synthetic', 'config': config,
            'e2e': {'value': n_forecasts * Ke / dt, 'unit': 'forecast-steps/s',
                    'h2d_bytes_per_step': h2d / Ke, 'd2h_bytes_per_step': d2h, 'api': api,
                    'steps': Ke, 'seconds': dt},
            'gpu_launches': launches, 'clocks': clocks.summary(), 'roofline': roofline, 'cpu_baseline': cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=256, help='forecasts (samples) per GPU')
    ap.add_argument('--e2e-steps', type=int, default=40, help='steps of the host-API measurement (<= --steps)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--math', default=os.environ.get('DLWP_MATH', 'tc'), choices=['tc', 'ffma'],
                    help='tc: tcgen05 tensor cores, fp16 hi/lo split x3 (fp32-level accuracy); ffma: fp32 FFMA2 kernels')
    ap.add_argument('--parallel', default='latband', choices=['latband', 'batch'],
                    help='N>1: latitude bands + halo exchange (north star) or independent forecasts per GPU')
    ap.add_argument('--halo', default='auto', choices=['auto', 'p2p', 'nccl'],
                    help='latitude bands: halo rows over peer memory (the conv epilogue stores them into the neighbours) or '
                         'one grouped NCCL SendRecv per step; auto = p2p when the plan qualifies')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='latitude bands: weak = --batch forecasts per GPU in flight (global batch grows with N), '
                         'strong = --batch forecasts in total')
    ap.add_argument('--workload', default='net_a', choices=['net_a', 'net_b'],
                    help='net_a: BASELINE.json configs[1] (the headline); net_b: the skip U-Net of configs[2-3]')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'bf16'],
                    help='tensor-core chain arithmetic: fp32-equivalent (fp16 hi/lo split x3) or plain bf16 (configs[2])')
    args = ap.parse_args()
    os.environ['DLWP_MATH'] = args.math
    if args.precision == 'bf16':
        os.environ['DLWP_PRECISION'] = 'bf16'
    if args.workload == 'net_b' and args.batch == 256:
        args.batch = 16
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit('launch multi-GPU runs with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d '
                         '--master-addr 127.0.0.1 --master-port 29500 bench.py --gpus %d ...' % (args.gpus, args.gpus))
    if args.workload == 'net_b':
        run_net_b(args, rank, world, local_rank)
    elif world > 1 and args.parallel == 'latband':
        run_latband(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
