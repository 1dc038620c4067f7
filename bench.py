#!/usr/bin/env python
"""bench.py -- the fractional-operator hot path on B200, one JSON line on rank 0.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "cfg 2"): growing-history 'caputo' fractional derivative of one
smooth synthetic field on N_NODES = 1e5 nodes, FP64.  One step = weight-table generation (K1) + field
padding + history sum (K2 main + fix-up) for `--alphas-per-gpu` values of alpha per GPU (default 1).
With N > 1 GPUs the alphas form a sweep dealt in contiguous blocks to the ranks (cfg 5 pattern, weak
scaling) and every step ends with an NCCL all_gather of the per-rank results.
Metric: node-evals/s (one node-eval = one output y_i, i.e. one full history row).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_NODES = 100000
CFG3_NODES, CFG3_FIELDS = 16384, 4096
RULE = 'caputo'
METRIC = 'fractional_history_node_evals_per_sec'
UNIT = 'node-evals/s'
FLUSH_BYTES = 256 << 20      # > 126 MB L2


def smooth_field_np(n, h):
    import numpy as np
    x = np.arange(n, dtype=np.float64) * h
    return 3.0 * np.cos(3.0 * x) + 2.0 * x      # g = f' for f = sin 3x + x^2 (SURVEY.md 8(d) cfg 2)


def workload_alphas(n_total):
    import numpy as np
    return np.array([0.5]) if n_total == 1 else np.linspace(0.1, 0.9, n_total)


# --------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, power, reasons = [], None, [], set()
        for t, line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9 or not (t_begin - 0.05 <= t <= t_end + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm), 'power_w_max': max(power) if power else None}


# --------------------------------------------------------------------------------------------------
# CPU legs: the oracle (plain-C restatement of the reference's algorithm: every row rebuilds its stencil
# with O(i) pow calls, then sums), all host threads, on a bounded row sample, extrapolated by sum(i+1)
# --------------------------------------------------------------------------------------------------
def cpu_reference_leg(budget_s):
    import numpy as np
    from oracle import clib
    h = 1.0 / N_NODES
    g = smooth_field_np(N_NODES + 2, h)
    cores = len(os.sched_getaffinity(0))     # all host threads we may use (torchrun pins OMP_NUM_THREADS=1: override)
    # calibrate on a thin sample, then size the stride for the time budget
    rows = np.arange(4096, N_NODES + 1, 4096, dtype=np.int64)
    t0 = time.perf_counter()
    clib.history_rows_naive(RULE, 0.5, h, g, rows, n_threads=cores)
    t_cal = time.perf_counter() - t0
    w_cal = float(np.sum(rows + 1.0))
    total_w = N_NODES * (N_NODES + 3) / 2.0
    per_weight = t_cal / w_cal
    stride = int(max(1, math.ceil(total_w * per_weight / budget_s)))
    rows = np.arange(stride, N_NODES + 1, stride, dtype=np.int64)
    if rows[-1] != N_NODES:
        rows = np.append(rows, N_NODES)
    t0 = time.perf_counter()
    y = clib.history_rows_naive(RULE, 0.5, h, g, rows, n_threads=cores)
    dt = time.perf_counter() - t0
    t_full = dt * total_w / float(np.sum(rows + 1.0))
    assert np.all(np.isfinite(y))
    return {'value': N_NODES / t_full, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'oracle/frx_oracle.c orc_history_rows_naive (reference algorithm: stencil of Settings(alpha,i*h,i) '
                      'rebuilt per row), OpenMP %d threads, rows i = %d*k (%d rows, %.1f s), full N=1e5 job extrapolated by '
                      'sum(i+1)' % (cores, stride, len(rows), dt),
            'seconds_sample': dt, 'seconds_full_extrapolated': t_full}


def python_reference_leg():
    """The same path in pure Python (oracle/ref_port.py, bit-identical to the unmodified reference):
    what a user of the reference actually waits for.  Three rows, one core, extrapolated."""
    from oracle import ref_port
    h = 1.0 / N_NODES
    g = list(smooth_field_np(N_NODES + 2, h))
    rows = (20000, 40000)
    t0 = time.perf_counter()
    ref_port.history(RULE, 0.5, h, g, rows)
    dt = time.perf_counter() - t0
    t_full = dt * (N_NODES * (N_NODES + 3) / 2.0) / sum(i + 1.0 for i in rows)
    return {'value': N_NODES / t_full, 'unit': UNIT, 'cores': 1, 'seconds_full_extrapolated': t_full}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    steps, warmup = args.steps, args.warmup
    budget = max(0.5, min(4.0, 150.0 / (steps + warmup)))
    vals = []
    last = None
    for k in range(warmup + steps):
        last = cpu_reference_leg(budget)
        if k >= warmup:
            vals.append(last['seconds_full_extrapolated'])
    t_full = sum(vals) / len(vals)
    value = N_NODES / t_full
    last['value'] = value
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warmup, 'ms_per_step': t_full * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(1, 1),
        'cpu_baseline': last,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'the reference is pure Python with an absent dependency (fdm); this arm times the C port of its '
                'algorithm (bit-identical results, tests/test_oracle_golden.py) on all host threads; each step is a '
                'bounded row sample extrapolated to the full N=1e5 job',
    }
    print(json.dumps(line))
    return 0


def workload_config(world, alphas_per_gpu, workload='cfg2'):
    if workload == 'cfg3':
        return {
            'workload': 'cfg3: batched Toeplitz application, N=16384 nodes x %d fields per GPU (smooth family '
                        "g_b = d/dx[sin((1+b/B)3x) + x^2]), rule 'caputo', alpha=0.5, FP64 tensor cores (DMMA); step = K1 weight "
                        'table + K3 application' % CFG3_FIELDS,
            'n_nodes': CFG3_NODES, 'n_fields': CFG3_FIELDS, 'rule': RULE,
            'l2': 'inputs (%.0f MB per GPU) larger than L2; no flush needed' % (CFG3_FIELDS * (CFG3_NODES + 2) * 8 / 1e6),
            'parallelism': 'independent fields per GPU' if world > 1 else 'single GPU',
        }
    return {
        'workload': 'cfg2: 1D fractional derivative history sum, N=1e5 nodes, single smooth field '
                    "g=3cos3x+2x, rule 'caputo', FP64; %d alpha per GPU (%s); step = K1 weight tables + K2 history sum%s"
                    % (alphas_per_gpu, 'alpha=0.5' if world * alphas_per_gpu == 1 else
                       'alpha sweep linspace(0.1,0.9,%d) in contiguous blocks per rank' % (world * alphas_per_gpu),
                       '' if world == 1 else ' + gather of all results on every GPU (see config.gather)'),
        'n_nodes': N_NODES, 'n_fields': 1, 'alphas_per_gpu': alphas_per_gpu, 'rule': RULE,
        'l2': 'flushed between timed steps by writing a %d MiB buffer' % (FLUSH_BYTES >> 20),
        'parallelism': 'independent alpha per GPU' if world > 1 else 'single GPU',
    }


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fractulus_b200 import _lib, batched, sweep

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- fractulus_b200 has no CPU path')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if args.workload == 'cfg3':
        return run_b200_cfg3(args, world, rank, local_rank, dev)
    if args.workload == 'cfg4':
        return run_b200_cfg4(args, world, rank, local_rank, dev)
    A = args.alphas_per_gpu
    steps, warmup = args.steps, max(args.warmup, 3)
    h = 1.0 / N_NODES
    alphas = workload_alphas(world * A)
    b, e = sweep.partition(len(alphas), world, rank)
    my_alphas = torch.as_tensor(alphas[b:e], device=dev)
    g_host = torch.from_numpy(smooth_field_np(N_NODES + 2, h)).pin_memory()
    g = g_host.to(dev)
    y = torch.empty((A, 1, N_NODES + 1), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * A, N_NODES + 1), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    ctx = _lib.context(local_rank)

    # N > 1: the gather is fused into the history kernel (its epilogue stores every row into all peers' copies of
    # the gathered array through symmetric-memory peer pointers; a device barrier ends the step).  NCCL all_gather
    # is the fallback when symmetric memory cannot be set up (or with --nccl-gather).
    peer, gather_mode = None, 'none'
    if world > 1:
        gather_mode = 'nccl all_gather'
        if not args.nccl_gather:
            ok = 1
            try:
                peer = sweep.PeerGather(world * A, N_NODES + 1)
            except Exception as ex:      # noqa: BLE001 -- any failure means: use NCCL
                ok, peer = 0, None
                sys.stderr.write('PeerGather unavailable on rank %d: %r\n' % (rank, ex))
            t = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if t.item() == 0:
                peer = None
            else:
                gather_mode = 'fused P2P stores from the kernel epilogue + symmetric-memory barrier'

    def step():
        if peer is not None:
            batched.history_gather(RULE, my_alphas, h, g, peer.ptrs, b, N_NODES)
            peer.barrier()
            return
        batched.history(RULE, my_alphas, h, g, N=N_NODES, out=y, validate=False)
        if world > 1:
            dist.all_gather_into_tensor(gathered, y.view(A, N_NODES + 1))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_load0 = time.time()
    # warm-up: ~1 s of the compute part alone (rank-local loop, so no collective inside: every rank may run a
    # different number of iterations) so that the sampled clocks are clocks under load, then W full steps
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 1.0:
        batched.history(RULE, my_alphas, h, g, N=N_NODES, out=y, validate=False)
        n += 1
        if n % 16 == 0:
            torch.cuda.synchronize()
    barrier()
    for _ in range(warmup):
        step()
    barrier()

    # ---- timed region: exactly K steps, device time per step, L2 flushed between steps -------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches0 = ctx.launch_count()
    barrier()
    wall0 = time.perf_counter()
    for k in range(steps):
        flush.zero_()
        ev[k][0].record()
        step()
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count() - launches0
    ms = [a.elapsed_time(b_) for a, b_ in ev]
    ms_per_step = sum(ms) / steps
    if world > 1:
        t = torch.tensor([ms_per_step], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = t.item()
    value = world * A * N_NODES / (ms_per_step * 1e-3)

    # ---- per-kernel pass: same K steps with CUDA events around every kernel (roofline numbers) -----
    ctx.set_timing(True)
    for k in range(steps):
        flush.zero_()
        batched.history(RULE, my_alphas, h, g, N=N_NODES, out=y, validate=False)
    ktimes = ctx.kernel_times()
    ctx.set_timing(False)
    t_load1 = time.time()
    clocks = sampler.stop(t_load0, t_load1) if rank == 0 else None

    # ---- e2e: host buffers in, host buffers out, through the public host-buffer API ---------------
    g_np = g_host.numpy()
    if world == 1:
        y_pinned = batched.pinned_empty((A, 1, N_NODES + 1))      # pinned input AND output: the kernels use them in place

        def e2e_step():
            return batched.history_host(RULE, alphas[b:e], h, g_np, N=N_NODES, device=local_rank, out=y_pinned)
        h2d = 8 * (N_NODES + 2) + 8 * A
        d2h = 8 * (N_NODES + 1) * A
    else:
        def e2e_step():
            return sweep.history_sweep(RULE, alphas, h, g_np, N=N_NODES, host_result='root')
        h2d = 8 * (N_NODES + 2)
        d2h = 8 * (N_NODES + 1) * A * world     # on rank 0 (the gathered result); 0 on the other ranks
    for _ in range(3):
        y_e2e = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        y_e2e = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_value = world * A * N_NODES / e2e_s
    # the host-buffer path and the device path give the same bits
    if peer is not None:     # the timed path wrote into the gathered array only
        torch.cuda.synchronize()
        y_dev = peer.buffer[b:e].cpu().numpy().reshape(A, N_NODES + 1)
    else:
        y_dev = y.cpu().numpy().reshape(A, N_NODES + 1)
    same = None
    if y_e2e is not None:
        same = bool(np.array_equal(np.asarray(y_e2e).reshape(-1, N_NODES + 1)[b:e], y_dev, equal_nan=True))

    if rank == 0:
        peak_dfma, peak_dmma = ctx.fp64_peak()
        main_ms, main_calls = ktimes['history_main']
        main_avg_ms = main_ms / max(main_calls, 1)
        algo_flop = A * (N_NODES * (N_NODES + 1) / 2.0) * 2.0        # N(N+1)/2 FMA per (alpha, field)
        achieved = algo_flop / (main_avg_ms * 1e-3) / 1e12
        per_step_kernel_ms = {k: v[0] / steps for k, v in ktimes.items() if v[1]}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        algo_bytes = A * 8.0 * (N_NODES + 1) + 8.0 * (N_NODES + 2)
        roofline = {
            'kernel': 'frx_k2_history_main (FP64 tensor cores, mma.m8n8k4.f64 = SASS DMMA)', 'bound': 'tensor',
            'achieved': achieved, 'peak': peak_dmma,
            'unit': 'TFLOP/s', 'frac': achieved / peak_dmma if peak_dmma else None,
            'traffic': K2_DRAM_TRAFFIC_BYTES * A if A == 1 else None,
            'peak_source': 'measured in this run: register-only DMMA m8n8k4 f64 probe (frx_measure_fp64_peak); '
                           'MEASURED_PEAKS.json has no FP64 figure (bf16/HBM only); the DFMA-pipe probe reads %.2f TFLOP/s '
                           '(frac of that: %.3f)' % (peak_dfma, achieved / peak_dfma if peak_dfma else float('nan')),
            'algorithmic_flop_per_launch': algo_flop, 'algorithmic_bytes_per_launch': algo_bytes,
            'kernel_ms': main_avg_ms, 'kernel_share_of_step': main_avg_ms / sum(per_step_kernel_ms.values()),
            'per_step_kernel_ms': per_step_kernel_ms,
            'hbm_view': {'achieved_gbs': algo_bytes / (main_avg_ms * 1e-3) / 1e9, 'peak_gbs': peaks.get('hbm_gbs'),
                         'note': 'intensity ~6e3 flop/B: not HBM bound'},
            'timing': 'per-kernel CUDA events on the launching stream, separate pass of the same K steps',
        }
        cpu = cpu_reference_leg(8.0) if not args.skip_cpu else {'skipped': '--skip-cpu (tuning run, not a bench line)'}
        try:
            if args.skip_cpu:
                raise RuntimeError('skipped')
            cpu['python_port'] = python_reference_leg()
        except Exception as ex:      # never let the informational leg break the line
            cpu['python_port'] = {'error': repr(ex)}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': dict(workload_config(world, A), gather=gather_mode),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_s * 1e3, 'api': 'frx_history_host (C ABI, pinned host buffers in and out; the prepare kernel reads the field and the history kernel stores the rows over PCIe in place)' if world == 1 else
                    'fractulus_b200.sweep.history_sweep (host field in on every rank, gathered host result out on rank 0)',
                    'bitwise_equal_to_device_path': same},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
            'wall_s_timed_region': wall, 'ms_per_step_min': min(ms), 'ms_per_step_max': max(ms),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0



def run_b200_cfg3(args, world, rank, local_rank, dev):
    """BASELINE cfg 3 (not the headline line; `--workload cfg3`): N = 16384 x 4096 fields per GPU through K3."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from fractulus_b200 import _lib, batched
    N, B = CFG3_NODES, CFG3_FIELDS
    steps, warmup = args.steps, max(args.warmup, 3)
    h = 1.0 / N
    x = torch.arange(N + 2, dtype=torch.float64, device=dev) * h
    k = 3.0 * (1.0 + (torch.arange(B, dtype=torch.float64, device=dev) + rank * B) / (B * world))
    G = k[:, None] * torch.cos(k[:, None] * x[None, :]) + 2.0 * x[None, :]
    Y = torch.empty((B, N + 1), dtype=torch.float64, device=dev)
    ctx = _lib.context(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_load0 = time.time()
    for _ in range(warmup):
        batched.toeplitz_apply(RULE, 0.5, h, G, N=N, out=Y)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    ev0.record()
    for _ in range(steps):
        batched.toeplitz_apply(RULE, 0.5, h, G, N=N, out=Y)
    ev1.record()
    barrier()
    launches = ctx.launch_count() - launches0
    ms_per_step = ev0.elapsed_time(ev1) / steps
    if world > 1:
        t = torch.tensor([ms_per_step], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = t.item()
    ctx.set_timing(True)
    for _ in range(steps):
        batched.toeplitz_apply(RULE, 0.5, h, G, N=N, out=Y)
    ktimes = ctx.kernel_times()
    ctx.set_timing(False)
    clocks = sampler.stop(t_load0, time.time()) if rank == 0 else None
    if rank == 0:
        peak_dfma, peak_dmma = ctx.fp64_peak()
        main_ms = ktimes['history_main'][0] / max(ktimes['history_main'][1], 1)
        flop = float(N) * (N + 1) * B
        achieved = flop / (main_ms * 1e-3) / 1e12
        line = {
            'metric': METRIC, 'value': world * B * N / (ms_per_step * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': steps,
            'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(world, 1, 'cfg3'), 'gpu_launches': launches,
            'clocks': clocks,
            'roofline': {'kernel': 'frx_k2_history_main (tensor variant, fields as problems)', 'bound': 'tensor',
                         'achieved': achieved, 'peak': peak_dmma, 'unit': 'TFLOP/s', 'frac': achieved / peak_dmma,
                         'traffic': None, 'peak_source': 'measured in this run: register-only DMMA m8n8k4 probe '
                         '(frx_measure_fp64_peak); DFMA probe %.2f TFLOP/s' % peak_dfma,
                         'algorithmic_flop_per_launch': flop, 'kernel_ms': main_ms,
                         'per_step_kernel_ms': {kk: v[0] / steps for kk, v in ktimes.items() if v[1]}},
            'e2e': None, 'cpu_baseline': None,
            'note': 'secondary workload (BASELINE configs[2]); the headline line is the default cfg2 run',
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_b200_cfg4(args, world, rank, local_rank, dev):
    """BASELINE cfg 4 (`--workload cfg4 --solve-n {64,128,256}`): assemble + batch-solve 1e5 systems per GPU over a
    400 alpha x 250 length-scale sweep.  Metric: solves/s."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from fractulus_b200 import _lib, batched
    n, n_alpha, n_ls = args.solve_n, 400, 250
    n_sys = n_alpha * n_ls
    steps, warmup = args.steps, max(args.warmup, 3)
    h = 1.0 / (n + 1)
    alpha = np.repeat(np.linspace(0.05, 0.95, n_alpha), n_ls)
    res = np.tile(1.0 + (np.arange(n_ls) % 30), n_alpha)
    xg = torch.arange(1, n + 1, dtype=torch.float64, device=dev) * h
    rhs = torch.sin(math.pi * xg).repeat(n_sys, 1).contiguous()
    ctx = _lib.context(local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_load0 = time.time()
    for _ in range(warmup):
        x = batched.assemble_solve(RULE, alpha, h, res, rhs)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    ev0.record()
    for _ in range(steps):
        x = batched.assemble_solve(RULE, alpha, h, res, rhs)
    ev1.record()
    torch.cuda.synchronize()
    launches = ctx.launch_count() - launches0
    ms_per_step = ev0.elapsed_time(ev1) / steps
    if world > 1:
        t = torch.tensor([ms_per_step], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = t.item()
    ctx.set_timing(True)
    for _ in range(steps):
        x = batched.assemble_solve(RULE, alpha, h, res, rhs)
    ktimes = ctx.kernel_times()
    ctx.set_timing(False)
    clocks = sampler.stop(t_load0, time.time()) if rank == 0 else None
    if rank == 0:
        peak_dfma, _ = ctx.fp64_peak()
        solve_ms = ktimes['solve'][0] / max(ktimes['solve'][1], 1)
        # flops the band-limited elimination needs: per column k, nb multipliers + nb x (2nb + 1) updates (x2) + b
        nb = np.minimum(res + 1, n - 1)
        flop = float(np.sum(n * (2.0 * nb * (2 * nb + 1) + 3 * nb) + n * (4 * nb + 1)))
        dense_flop = n_sys * (2.0 / 3.0 * n ** 3 + 2.0 * n * n)
        achieved = flop / (solve_ms * 1e-3) / 1e12
        line = {
            'metric': 'fractional_bvp_solves_per_sec', 'value': world * n_sys / (ms_per_step * 1e-3), 'unit': 'solves/s',
            'n_gpus': world, 'steps': steps, 'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'cfg4: assemble + batch-solve %d systems per GPU, n=%d unknowns, 400 alpha in [0.05,0.95] x 250 '
                                   "length scales (resolution 1..30), rule 'caputo' Riesz-Caputo operator, rhs sin(pi x); step = K5 "
                                   'stencil weights + K4 in-kernel assembly + LU' % (n_sys, n), 'n_systems': n_sys, 'n': n},
            'gpu_launches': launches, 'clocks': clocks,
            'roofline': {'kernel': 'frx_k4_assemble_solve', 'bound': 'fp64', 'achieved': achieved, 'peak': peak_dfma,
                         'unit': 'TFLOP/s', 'frac': achieved / peak_dfma, 'traffic': None, 'kernel_ms': solve_ms,
                         'algorithmic_flop_per_launch': flop, 'dense_lu_flop_for_context': dense_flop,
                         'peak_source': 'measured in this run: register-only DFMA probe',
                         'per_step_kernel_ms': {kk: v[0] / steps for kk, v in ktimes.items() if v[1]},
                         'note': 'band-limited LU: flop count is what the banded elimination needs, not 2/3 n^3'},
            'e2e': None, 'cpu_baseline': None,
            'note': 'secondary workload (BASELINE configs[3]); parity unpinned (no reference solver)',
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0

_JSON_LINES = []


def emit(line):
    _JSON_LINES.append(json.dumps(line))


# dram__bytes_read.sum + dram__bytes_write.sum of frx_k2_history_main per launch (cfg 2, one alpha) from the committed
# `ncu --set full` capture profiles/r01_k2_history_main_v6_final_ncu_full.txt: 3.27 MB read (weight tables + padded
# field; the algorithmic 1.6 MB are the field in and y out), 0 B written back during the kernel (y stays in L2).
K2_DRAM_TRAFFIC_BYTES = 3266816


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=('b200', 'reference'))
    ap.add_argument('--alphas-per-gpu', type=int, default=1)
    ap.add_argument('--workload', default='cfg2', choices=('cfg2', 'cfg3', 'cfg4'),
                    help='cfg2 (default, the headline line; with --alphas-per-gpu 1250 it is the cfg5 sweep) or cfg3')
    ap.add_argument('--solve-n', type=int, default=128, help='cfg4: unknowns per system (64..256)')
    ap.add_argument('--nccl-gather', action='store_true', help='N > 1: gather with NCCL all_gather instead of the fused path')
    ap.add_argument('--skip-cpu', action='store_true', help='tuning only: skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    # stdout carries exactly one JSON line: libraries that chat on fd 1 (NCCL prints its version banner there)
    # are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        return run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
        if _JSON_LINES:
            sys.stdout.write('\n'.join(_JSON_LINES) + '\n')
            sys.stdout.flush()


if __name__ == '__main__':
    sys.exit(main())
