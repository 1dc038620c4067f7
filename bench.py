#!/usr/bin/env python
"""bench.py -- the fractional-operator hot path on B200, one JSON line on rank 0.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Headline (BASELINE.json configs[1], "cfg 2"): growing-history 'caputo' fractional derivative of one smooth synthetic field
on N_NODES = 1e5 nodes, FP64.  One step = weight tables (K0/K1) + field padding + history sum (K2) for `--alphas-per-gpu`
values of alpha per GPU: 1 at N=1 (cfg 2 itself); at N>1 the default is 1250 alpha per GPU, i.e. the cfg-5 sweep slice
(1e4 alpha on 8 GPUs), dealt in contiguous blocks, every step ending with the gather of all rows onto every GPU -- fused
into the history kernel (P2P stores over NVLink from its epilogue + a device barrier).
Metric: node-evals/s (one node-eval = one output y_i, i.e. one full history row).

The same line carries one sub-record per remaining BASELINE config under "configs" (each with ms_per_step, roofline,
e2e with host buffers, cpu_baseline, a parity spot check): cfg1 (reference API, N=1000), cfg3 (16384 x 4096 fields),
cfg4 (1e5 assemble+solves, n = 64 / 128 / 256), cfg5_slice (1250 alpha on one GPU; at --gpus 8 the headline IS cfg 5),
and the latency of the drop-in's small calls.  `--skip-configs` drops them (tuning runs).
"""
import argparse
import glob
import json
import math
import os
import re
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
CFG5_ALPHAS_PER_GPU = 1250
RULE = 'caputo'
METRIC = 'fractional_history_node_evals_per_sec'
UNIT = 'node-evals/s'
FLUSH_BYTES = 256 << 20      # > 126 MB L2
FUSED_GATHER = 'fused P2P stores from the kernel epilogue + symmetric-memory barrier'


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

    def window(self, t_begin, t_end):
        sm, smax, power, reasons = [], None, [], set()
        for t, line in list(self.lines):
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

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        out = self.window(t_begin, t_end)
        self.proc.terminate()
        return out


# --------------------------------------------------------------------------------------------------
# CPU legs.  The reference itself (unmodified copy in baseline/_ref + the fdm shim, oracle/ref_runner.py) when that copy
# exists, on all host cores; else the plain-C port of its algorithm (oracle/frx_oracle.c).  A step of the reference
# arm is a BOUNDED SAMPLE of the cfg-2 job: evenly spaced rows i = stride*k (mean row length N/2, like the full job),
# value = rows of the sample / wall time of the sample -- measured, not extrapolated.
# --------------------------------------------------------------------------------------------------
class ReferenceArm(object):
    def __init__(self):
        from oracle import ref_runner
        self.cores = len(os.sched_getaffinity(0))     # all host threads we may use (torchrun pins OMP_NUM_THREADS=1: override)
        self.h = 1.0 / N_NODES
        self.kind = 'reference' if ref_runner.available() else 'port'
        self.pool = ref_runner.RowPool(N_NODES, self.h, self.cores) if self.kind == 'reference' else None
        self.per_row_s = None

    def sample_rows(self, budget_s):
        import numpy as np
        if self.per_row_s is None:      # calibrate: seconds of one core per weight
            rows = np.linspace(N_NODES // (2 * self.cores), N_NODES, 2 * self.cores).astype(np.int64)
            dt, _ = self.run(rows)
            self.per_row_s = dt * self.cores / float(np.sum(rows + 1.0))
        n_rows = int(max(self.cores, min(N_NODES, budget_s * self.cores / (self.per_row_s * (N_NODES / 2.0)))))
        n_rows = max(self.cores, (n_rows // self.cores) * self.cores)
        stride = max(1, N_NODES // n_rows)
        rows = np.arange(stride, N_NODES + 1, stride, dtype=np.int64)
        return rows, stride

    def run(self, rows):
        import numpy as np
        if self.kind == 'reference':
            dt, ys = self.pool.rows(RULE, 0.5, [int(i) for i in rows])
            assert all(math.isfinite(v) for v in ys.values())
            return dt, ys
        from oracle import clib
        g = smooth_field_np(N_NODES + 2, self.h)
        t0 = time.perf_counter()
        y = clib.history_rows_naive(RULE, 0.5, self.h, g, np.asarray(rows, dtype=np.int64), n_threads=self.cores)
        dt = time.perf_counter() - t0
        assert np.all(np.isfinite(y))
        return dt, dict(zip([int(i) for i in rows], [float(v) for v in y]))

    def step(self, budget_s):
        rows, stride = self.sample_rows(budget_s)
        dt, ys = self.run(rows)
        what = ('the UNMODIFIED reference (baseline/_ref: fractulus.create_left_stencil(\'caputo\', Settings(alpha, i*h, i)) + '
                'sum(g(node)*weight) per row, reference test/integration/test_equation.py:51-57) on oracle/fdm_shim.py, '
                'multiprocessing' if self.kind == 'reference' else
                'oracle/frx_oracle.c orc_history_rows_naive (C port of the reference algorithm, stencil rebuilt per row), OpenMP')
        return {'value': len(rows) / dt, 'unit': UNIT, 'cores': self.cores, 'kind': self.kind, 'seconds_sample': dt,
                'rows_in_sample': int(len(rows)),
                'sample': '%s, %d processes/threads; rows i = %d*k, k = 1..%d of the N=1e5 cfg-2 job (mean row length N/2 like the '
                          'full job), value = rows / seconds of the sample' % (what, self.cores, stride, len(rows)),
                'seconds_full_job_at_this_rate': N_NODES * dt / len(rows)}, ys

    def close(self):
        if self.pool is not None:
            self.pool.close()


def c_port_rate(budget_s=3.0):
    """The plain-C port (OpenMP, all threads) on the same kind of sample: context next to the reference's own rate."""
    import numpy as np
    from oracle import clib
    cores = len(os.sched_getaffinity(0))
    h = 1.0 / N_NODES
    g = smooth_field_np(N_NODES + 2, h)
    rows = np.arange(2048, N_NODES + 1, 2048, dtype=np.int64)
    t0 = time.perf_counter()
    clib.history_rows_naive(RULE, 0.5, h, g, rows, n_threads=cores)
    dt = time.perf_counter() - t0
    stride = int(max(1, math.ceil(2048 * dt / budget_s)))
    rows = np.arange(stride, N_NODES + 1, stride, dtype=np.int64)
    t0 = time.perf_counter()
    clib.history_rows_naive(RULE, 0.5, h, g, rows, n_threads=cores)
    dt = time.perf_counter() - t0
    return {'value': len(rows) / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'oracle/frx_oracle.c, OpenMP %d threads, rows i = %d*k (%d rows, %.1f s)' % (cores, stride, len(rows), dt)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    steps, warmup = args.steps, args.warmup
    budget = max(0.5, min(6.0, 150.0 / (steps + warmup)))
    arm = ReferenceArm()
    secs, rows_n, last = [], [], None
    for k in range(warmup + steps):
        last, _ = arm.step(budget)
        if k >= warmup:
            secs.append(last['seconds_sample'])
            rows_n.append(last['rows_in_sample'])
    arm.close()
    value = sum(rows_n) / sum(secs)
    last['value'] = value
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warmup, 'ms_per_step': 1e3 * sum(secs) / len(secs), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        # the same config dict as the B200 arm prints for this --gpus (the CPU arm itself always samples the single-field job:
        # every alpha of the sweep costs the same)
        'config': workload_config(args.gpus, 1 if args.gpus == 1 else CFG5_ALPHAS_PER_GPU,
                                  gather='none' if args.gpus == 1 else FUSED_GATHER),
        'cpu_baseline': last,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'CPU arm: each step is a bounded sample of the cfg-2 job (evenly spaced rows), timed on the host cores; '
                'kind "reference" = the unmodified reference package itself (pure Python, its fdm dependency replaced by the '
                'repo\'s shim), kind "port" = the C restatement when baseline/_ref is absent',
    }
    print(json.dumps(line))
    return 0


def workload_config(world, alphas_per_gpu, gather='none'):
    total = world * alphas_per_gpu
    return {
        'workload': 'cfg2: 1D fractional derivative history sum, N=1e5 nodes, single smooth field '
                    "g=3cos3x+2x, rule 'caputo', FP64; %d alpha per GPU (%s); step = K1 weight tables + K2 history sum%s"
                    % (alphas_per_gpu, 'alpha=0.5' if total == 1 else
                       'alpha sweep linspace(0.1,0.9,%d) in contiguous blocks per rank%s'
                       % (total, ' = the cfg5 sweep (1e4 alpha x N=1e5) at 8 GPUs' if alphas_per_gpu == CFG5_ALPHAS_PER_GPU else ''),
                       '' if world == 1 else ' + gather of all results on every GPU (see config.gather)'),
        'n_nodes': N_NODES, 'n_fields': 1, 'alphas_per_gpu': alphas_per_gpu, 'rule': RULE,
        'l2': 'flushed between timed steps by writing a %d MiB buffer' % (FLUSH_BYTES >> 20),
        'parallelism': 'independent alpha per GPU' if world > 1 else 'single GPU', 'gather': gather,
    }


def newest_profile_traffic(pattern):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest committed `ncu --set full` extract matching
    `pattern` under profiles/ (file name reported next to the number)."""
    files = sorted(glob.glob(os.path.join(REPO, 'profiles', pattern)))
    for path in reversed(files):
        total, found = 0.0, 0
        for ln in open(path):
            m = re.match(r'dram__bytes_(read|write)\.sum \[(\w+)\] = ([\d.eE+-]+)', ln)
            if m and found < 2:
                scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(m.group(2), 1.0)
                total += float(m.group(3)) * scale
                found += 1
        if found == 2:
            return total, os.path.relpath(path, REPO)
    return None, None


# --------------------------------------------------------------------------------------------------
# the history workload (headline at every N; cfg5_slice sub-record at N=1)
# --------------------------------------------------------------------------------------------------
def bench_history(env, A, steps, warmup, kernel_pass=True, e2e=True, warm_seconds=1.0):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fractulus_b200 import batched, sweep
    world, rank, dev, ctx = env['world'], env['rank'], env['dev'], env['ctx']
    h = 1.0 / N_NODES
    alphas = workload_alphas(world * A)
    b, e = sweep.partition(len(alphas), world, rank)
    my_alphas = torch.as_tensor(alphas[b:e], device=dev)
    g_host = torch.from_numpy(smooth_field_np(N_NODES + 2, h)).pin_memory()
    g = g_host.to(dev)
    y = torch.empty((A, 1, N_NODES + 1), dtype=torch.float64, device=dev) if world == 1 else None
    flush = env['flush']

    peer, gather_mode = None, 'none'
    if world > 1:
        gather_mode = 'nccl all_gather'
        gathered = None
        if not env['args'].nccl_gather:
            peer = sweep._peer_plan(world * A, N_NODES + 1, None, world)
        if peer is not None:
            gather_mode = FUSED_GATHER
        else:
            y = torch.empty((A, 1, N_NODES + 1), dtype=torch.float64, device=dev)
            gathered = torch.empty((world * A, N_NODES + 1), dtype=torch.float64, device=dev)

    def compute_only():
        if peer is not None:
            batched.history_gather(RULE, my_alphas, h, g, peer.ptrs[:1], b, N_NODES)      # own copy only: no NVLink traffic
        else:
            batched.history(RULE, my_alphas, h, g, N=N_NODES, out=y, validate=False)

    def step():
        if peer is not None:
            batched.history_gather(RULE, my_alphas, h, g, peer.ptrs, b, N_NODES)
            peer.barrier()
            return
        batched.history(RULE, my_alphas, h, g, N=N_NODES, out=y, validate=False)
        if world > 1:
            dist.all_gather_into_tensor(gathered, y.view(A, N_NODES + 1))

    barrier = env['barrier']
    t_load0 = time.time()
    # warm-up: ~1 s of the compute part alone (rank-local loop, no collective inside) so that the sampled clocks are
    # clocks under load, then W full steps
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < warm_seconds:
        compute_only()
        n += 1
        if n % 16 == 0 or A > 16:
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
    out = {'value': value, 'ms_per_step': ms_per_step, 'ms_per_step_min': min(ms), 'ms_per_step_max': max(ms),
           'gpu_launches': launches, 'wall_s_timed_region': wall, 'gather': gather_mode}

    # ---- single-GPU reference for the weak-scaling claim: the same per-GPU work without any gather ----
    if world > 1:
        k_steps = max(3, min(steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(k_steps):
            compute_only()
        e1.record()
        torch.cuda.synchronize()
        local_ms = e0.elapsed_time(e1) / k_steps
        t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out['same_work_no_gather'] = {'ms_per_step': t.item(), 'weak_scaling_efficiency_device': t.item() / ms_per_step,
                                      'note': 'max over ranks of this rank\'s %d alpha per step with the result stored to its '
                                              'own copy only (what one GPU alone does); efficiency = that / ms_per_step' % A}

    # ---- per-kernel pass: CUDA events around every kernel (roofline numbers) -----
    ktimes = None
    if kernel_pass:
        k_steps = steps if A == 1 else max(2, min(steps, 5))
        ctx.set_timing(True)
        for k in range(k_steps):
            flush.zero_()
            compute_only()
        ktimes = ctx.kernel_times()
        ctx.set_timing(False)
        out['kernel_pass_steps'] = k_steps
    out['t_load'] = (t_load0, time.time())

    # ---- e2e: host buffers in, host buffers out, through the public host-buffer API ---------------
    if e2e:
        g_np = g_host.numpy()
        e_steps = steps if A == 1 else max(2, min(steps, 5))
        if world == 1:
            y_pinned = batched.pinned_empty((A, 1, N_NODES + 1))      # pinned input AND output: the kernels use them in place

            def e2e_step():
                return batched.history_host(RULE, alphas[b:e], h, g_np, N=N_NODES, device=env['local_rank'], out=y_pinned)
            h2d = 8 * (N_NODES + 2) + 8 * A
            d2h = 8 * (N_NODES + 1) * A
            api = ('frx_history_host (C ABI, pinned host buffers in and out; the prepare kernel reads the field and the history '
                   'kernel stores the rows over PCIe in place)')
        else:
            def e2e_step():
                return sweep.history_sweep(RULE, alphas, h, g_np, N=N_NODES, host_result='root')
            h2d = (8 * (N_NODES + 2) + 8 * A) * world
            d2h = 8 * (N_NODES + 1) * A * world
            api = ('fractulus_b200.sweep.history_sweep (host field in on every rank; every rank\'s kernels store their rows in place '
                   'into ONE shared pinned host array over their own PCIe link; bytes are the totals over all ranks)')
        for _ in range(2):
            y_e2e = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            y_e2e = e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e_steps
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = t.item()
        # the host-buffer path and the device path give the same values (bitwise when the launch shape is the same)
        torch.cuda.synchronize()
        y_dev = (peer.buffer[b:e] if peer is not None else y.view(A, N_NODES + 1)).cpu().numpy().reshape(A, N_NODES + 1)
        same = None
        if y_e2e is not None:
            mine = np.asarray(y_e2e).reshape(-1, N_NODES + 1)[b:e]
            same = {'bitwise_equal_to_device_path': bool(np.array_equal(mine, y_dev, equal_nan=True)),
                    'max_rel_diff_to_device_path': float(np.nanmax(np.abs(mine[:, 1:] - y_dev[:, 1:]) / np.abs(y_dev[:, 1:])))}
        out['e2e'] = {'value': world * A * N_NODES / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                      'ms_per_step': e2e_s * 1e3, 'steps': e_steps, 'api': api, 'vs_device_path': same}
    out['_ktimes'], out['_alphas'], out['_b_e'] = ktimes, alphas, (b, e)
    out['_y_dev_rows'] = None
    # parity spot check on rank 0: a few rows of the first alpha against the correctly-rounded and the glibc oracle
    if rank == 0:
        from oracle import clib
        rows = np.array([1, 2, 1023, 1024, 1025, 50001, 77777, N_NODES - 1, N_NODES])
        y0 = (peer.buffer[b] if peer is not None else y.view(A, N_NODES + 1)[0]).cpu().numpy()
        gf = g_host.numpy()
        w_cr = clib.history_rows_naive(RULE, float(alphas[b]), h, gf, rows, cr=True)
        w_gl = clib.history_rows_naive(RULE, float(alphas[b]), h, gf, rows)
        out['parity_spot_check'] = {
            'rows': [int(r) for r in rows], 'alpha': float(alphas[b]),
            'max_rel_err_vs_correctly_rounded_oracle': float(np.max(np.abs(y0[rows] - w_cr) / np.abs(w_cr))),
            'max_rel_err_vs_reference_arithmetic': float(np.max(np.abs(y0[rows] - w_gl) / np.abs(w_gl))),
            'note': 'full-row checks: tests/test_gpu_history.py::test_cfg2_full_size_every_row_vs_both_oracles'}
    return out


def history_roofline(res, A, peak_dmma, peak_dfma, steps):
    ktimes = res['_ktimes']
    main_ms, main_calls = ktimes['history_main']
    main_avg_ms = main_ms / max(main_calls, 1)
    algo_flop = A * (N_NODES * (N_NODES + 1) / 2.0) * 2.0        # N(N+1)/2 FMA per (alpha, field)
    achieved = algo_flop / (main_avg_ms * 1e-3) / 1e12
    k_steps = res.get('kernel_pass_steps', steps)
    per_step_kernel_ms = {k: v[0] / k_steps for k, v in ktimes.items() if v[1]}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    algo_bytes = A * 8.0 * (N_NODES + 1) + 8.0 * (N_NODES + 2)
    traffic, traffic_file = newest_profile_traffic('r0*_k2_history_main*_final_ncu_full.txt') if A == 1 else (None, None)
    return {
        'kernel': 'frx_k2_history_main (FP64 tensor cores, mma.m8n8k4.f64 = SASS DMMA)', 'bound': 'tensor',
        'achieved': achieved, 'peak': peak_dmma, 'unit': 'TFLOP/s', 'frac': achieved / peak_dmma if peak_dmma else None,
        'traffic': traffic, 'traffic_source': traffic_file,
        'peak_source': 'measured in this run: register-only DMMA m8n8k4 f64 probe (frx_measure_fp64_peak); '
                       'MEASURED_PEAKS.json has no FP64 figure (bf16/HBM only); the DFMA-pipe probe reads %.2f TFLOP/s '
                       '(frac of that: %.3f)' % (peak_dfma, achieved / peak_dfma if peak_dfma else float('nan')),
        'algorithmic_flop_per_launch': algo_flop, 'algorithmic_bytes_per_launch': algo_bytes,
        'kernel_ms': main_avg_ms, 'kernel_share_of_step': main_avg_ms / sum(per_step_kernel_ms.values()),
        'per_step_kernel_ms': per_step_kernel_ms,
        'hbm_view': {'achieved_gbs': algo_bytes / (main_avg_ms * 1e-3) / 1e9, 'peak_gbs': peaks.get('hbm_gbs'),
                     'note': 'intensity ~6e3 flop/B: not HBM bound'},
        'timing': 'per-kernel CUDA events on the launching stream, separate pass of the same steps',
    }


def public(res):
    return {k: v for k, v in res.items() if not k.startswith('_') and k != 't_load'}


# --------------------------------------------------------------------------------------------------
# sub-records: the other BASELINE configs
# --------------------------------------------------------------------------------------------------
def sub_cfg1(env):
    """BASELINE configs[0]: D^0.5 of x^2 on N=1000 nodes through the reference's own API -- the unmodified reference timed on
    one core beside the drop-in (same call pattern, GPU weights) and the batched GPU path."""
    import numpy as np
    import torch
    import fractulus_b200 as fr
    from fractulus_b200 import batched, fdm_carrier
    from oracle import ref_runner
    N, h, alpha = 1000, 1e-3, 0.5
    out = {'workload': 'cfg1: Caputo derivative of x^2, alpha=0.5, N=1000 nodes; per node i: Operator(create_left_stencil(\'caputo\', '
                       'Settings(alpha, i*h, i)), Operator(Stencil.central(h))).expand(Point(i*h)), sum(f(node)*weight) '
                       '(reference test/integration/test_equation.py:51-66)', 'unit': UNIT}
    x = np.arange(N + 1) * h
    exact = math.gamma(3.) / math.gamma(2.5) * x[1:] ** 1.5
    ref_y = None
    if ref_runner.available():
        t_ref, ref_y = ref_runner.cfg1(alpha, N, h)
        ref_y = np.array(ref_y)
        out['cpu_baseline'] = {'value': N / t_ref, 'unit': UNIT, 'cores': 1, 'kind': 'reference', 'seconds': t_ref,
                               'sample': 'the whole job (N=1000) through the UNMODIFIED reference (baseline/_ref) on oracle/fdm_shim.py, '
                                         'one core (the reference is single-threaded)',
                               'max_rel_err_vs_analytic': float(np.max(np.abs(ref_y - exact) / exact))}
    else:
        out['cpu_baseline'] = {'unavailable': 'baseline/_ref missing (python -m oracle.make_ref needs /root/reference)'}
    # (a) the drop-in through the same API calls (one K5 launch + two copies per stencil, Python dict expansion)
    f = lambda pt: pt.x ** 2
    t0 = time.perf_counter()
    ys = []
    for i in range(1, N + 1):
        op = fdm_carrier.Operator(fr.create_left_stencil('caputo', fr.Settings(alpha, i * h, i)),
                                  fdm_carrier.Operator(fdm_carrier.Stencil.central(h)))
        ys.append(sum([f(node) * w for node, w in op.expand(fdm_carrier.Point(i * h)).items()]))
    t_api = time.perf_counter() - t0
    ys = np.array(ys)
    out['dropin_api'] = {'value': N / t_api, 'seconds': t_api, 'ms_per_step': t_api * 1e3,
                         'note': 'same 1000 API calls against fractulus_b200 (weights from kernel K5, host dict expansion)',
                         'max_rel_err_vs_reference': float(np.max(np.abs(ys - ref_y) / np.abs(ref_y))) if ref_y is not None else None,
                         'max_rel_err_vs_analytic': float(np.max(np.abs(ys - exact) / exact))}
    # (b) the batched GPU path: the inner central difference sampled into g, the outer operator = K1 + K2, host in / host out
    g = f_central = (x + h / 2.) ** 2 * (1. / h) + (x - h / 2.) ** 2 * (-1. / h)
    g_pin = batched.pinned_empty(g.shape)
    g_pin[:] = g
    y_pin = batched.pinned_empty((1, 1, N + 1))
    for _ in range(3):
        batched.history_host('caputo', [alpha], h, g_pin, N=N, device=env['local_rank'], out=y_pin)
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        yb = batched.history_host('caputo', [alpha], h, g_pin, N=N, device=env['local_rank'], out=y_pin)
    t_b = (time.perf_counter() - t0) / reps
    yb = np.array(yb).reshape(-1)[1:]
    out['value'], out['ms_per_step'] = N / t_b, t_b * 1e3
    out['e2e'] = {'value': N / t_b, 'unit': UNIT, 'ms_per_step': t_b * 1e3, 'h2d_bytes_per_step': 8 * (N + 1) + 8,
                  'd2h_bytes_per_step': 8 * (N + 1), 'api': 'batched.history_host (frx_history_host), pinned host buffers'}
    out['batched_path'] = {'note': 'one call: K0/K1 weight tables + K2 history sum for all 1000 rows, host buffers in and out',
                           'max_rel_err_vs_reference': float(np.max(np.abs(yb - ref_y) / np.abs(ref_y))) if ref_y is not None else None,
                           'max_rel_err_vs_analytic': float(np.max(np.abs(yb - exact) / exact))}
    out['roofline'] = {'bound': 'latency', 'note': 'N(N+1) = 1e6 flop: three kernel launches, ~30 us; no roofline applies at this size'}
    if 'seconds' in out['cpu_baseline']:
        out['speedup_vs_reference'] = {'batched_path': out['cpu_baseline']['seconds'] / t_b,
                                       'dropin_api_same_calls': out['cpu_baseline']['seconds'] / t_api}
    return out


def sub_small_calls(env):
    """Latency of the drop-in's create_left_stencil against the reference's (missing #6)."""
    import numpy as np
    import fractulus_b200 as fr
    from fractulus_b200 import batched
    from oracle import ref_runner
    points = [4, 100, 1000, 20000]
    out = {'workload': "create_left_stencil('caputo', Settings(0.5, p*0.01, p)) latency, best of 5 (seconds per call)", 'unit': 's/call'}
    gpu = {}
    for p in points:
        best = None
        for _ in range(5):
            t0 = time.perf_counter()
            fr.create_left_stencil('caputo', fr.Settings(0.5, p * 0.01, p))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        gpu[p] = best
    out['dropin_seconds_per_call'] = {str(p): gpu[p] for p in points}
    if ref_runner.available():
        ref = ref_runner.create_stencil_latency(points, repeats=3)
        out['reference_seconds_per_call'] = {str(p): ref[p] for p in points}
        out['reference_over_dropin'] = {str(p): ref[p] / gpu[p] for p in points}
    # the batched convenience: many Settings, one K5 launch, carriers back
    n = 2000
    settings = [fr.Settings(0.3 + 0.0002 * k, 0.04 * (4 + k % 13), 4 + k % 13) for k in range(n)]
    fr.create_left_stencils('caputo', settings[:8])
    t0 = time.perf_counter()
    sts = fr.create_left_stencils('caputo', settings)
    dt = time.perf_counter() - t0
    out['batched_create_left_stencils'] = {'n_settings': n, 'seconds_per_stencil': dt / n, 'total_nodes': int(sum(len(s.weights_array) for s in sts)),
                                          'note': 'fractulus_b200.create_left_stencils: one layout call + one K5 launch + one copy for the whole list'}
    return out


def sub_cfg3(env, steps=5):
    """BASELINE cfg 3: N = 16384 nodes x 4096 fields per GPU through K3 (one alpha, the DMMA dense-contraction path)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from fractulus_b200 import batched, sweep
    from oracle import clib
    world, rank, dev, ctx = env['world'], env['rank'], env['dev'], env['ctx']
    N, B = CFG3_NODES, CFG3_FIELDS
    h = 1.0 / N
    x = torch.arange(N + 2, dtype=torch.float64, device=dev) * h
    k = 3.0 * (1.0 + (torch.arange(B, dtype=torch.float64, device=dev) + rank * B) / (B * world))
    G = k[:, None] * torch.cos(k[:, None] * x[None, :]) + 2.0 * x[None, :]
    Y = torch.empty((B, N + 1), dtype=torch.float64, device=dev)
    for _ in range(3):
        batched.toeplitz_apply(RULE, 0.5, h, G, N=N, out=Y)
    env['barrier']()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    ev0.record()
    for _ in range(steps):
        batched.toeplitz_apply(RULE, 0.5, h, G, N=N, out=Y)
    ev1.record()
    env['barrier']()
    launches = ctx.launch_count() - launches0
    ms_per_step = ev0.elapsed_time(ev1) / steps
    if world > 1:
        t = torch.tensor([ms_per_step], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = t.item()
    ctx.set_timing(True)
    for _ in range(2):
        batched.toeplitz_apply(RULE, 0.5, h, G, N=N, out=Y)
    ktimes = ctx.kernel_times()
    ctx.set_timing(False)
    main_ms = ktimes['history_main'][0] / max(ktimes['history_main'][1], 1)
    flop = float(N) * (N + 1) * B
    achieved = flop / (main_ms * 1e-3) / 1e12
    out = {'workload': 'cfg3: batched Toeplitz application, N=16384 nodes x %d fields per GPU (smooth family g_b = d/dx[sin((1+b/B)3x) + x^2]), '
                       "rule 'caputo', alpha=0.5, FP64 tensor cores (DMMA); step = K1 weight table + K3 application; inputs (%.0f MB) larger "
                       'than L2' % (B, B * (N + 2) * 8 / 1e6),
           'value': world * B * N / (ms_per_step * 1e-3), 'unit': UNIT, 'ms_per_step': ms_per_step, 'steps': steps, 'gpu_launches': launches,
           'roofline': {'kernel': 'frx_k2_history_main (tensor variant, fields as problems)', 'bound': 'tensor', 'achieved': achieved,
                        'peak': env['peak_dmma'], 'unit': 'TFLOP/s', 'frac': achieved / env['peak_dmma'], 'traffic': None,
                        'algorithmic_flop_per_launch': flop, 'algorithmic_bytes_per_launch': 16.0 * N * B, 'kernel_ms': main_ms,
                        'per_step_kernel_ms': {kk: v[0] / 2 for kk, v in ktimes.items() if v[1]}}}
    # e2e: host fields in, host results out (pageable NumPy in; the shared pinned result of the sweep API out)
    Gh = batched.pinned_empty((B, N + 2))      # page-locked host fields (the e2e contract: inputs come from pinned host memory)
    torch.as_tensor(Gh).copy_(G)
    torch.cuda.synchronize()
    Yh_pinned = batched.pinned_empty((B, N + 1))
    if world == 1:
        def e2e_step():
            return sweep.toeplitz_apply_sweep(RULE, 0.5, h, Gh, N=N)
    else:      # weak scaling: every rank holds its own 4096 fields -- the per-rank host call the sweep API is made of
        def e2e_step():
            return batched.toeplitz_apply_host(RULE, 0.5, h, Gh, N=N, out=Yh_pinned)
    Yh = e2e_step()
    env['barrier']()
    t0 = time.perf_counter()
    for _ in range(2):
        Yh = e2e_step()
    env['barrier']()
    e2e_s = (time.perf_counter() - t0) / 2
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    out['e2e'] = {'value': world * B * N / e2e_s, 'unit': UNIT, 'ms_per_step': e2e_s * 1e3, 'h2d_bytes_per_step': 8 * B * (N + 2) * world,
                  'd2h_bytes_per_step': 8 * B * (N + 1) * world,
                  'api': 'sweep.toeplitz_apply_sweep (pinned NumPy fields in, shared pinned host result out; 512-field chunks, copies '
                         'overlapped with the kernel)' if world == 1 else
                         'batched.toeplitz_apply_host on every rank (its own block of fields; pinned NumPy in and out, chunked, overlapped)',
                  'note': 'PCIe-bound: %.0f MB each way per GPU' % (8 * B * (N + 1) / 1e6)}
    if rank == 0:
        f = 777
        gf = Gh[f]
        want = clib.history_full(RULE, 0.5, h, N, gf, cr=True)
        s_abs = clib.history_full(RULE, 0.5, h, N, gf, cr=True, abs_terms=True)
        err = np.abs(np.asarray(Yh)[f, 1:] - want[1:])
        out['parity_spot_check'] = {'field': f, 'rows': 'all %d' % N,
                                    'max_err_over_tolerance': float(np.max(err / (1e-12 * np.abs(want[1:]) + np.sqrt(np.arange(1, N + 1) / 4.) * np.finfo(float).eps * s_abs[1:]))),
                                    'tolerance': '1e-12 |y_i| + sqrt(i/4) eps sum_j |w_ij g_j| vs the correctly-rounded oracle (tests/test_gpu_history.py)'}
        # CPU baseline: the C port of the reference algorithm (no faster reference exists: the reference rebuilds each row's stencil)
        cores = len(os.sched_getaffinity(0))
        rows = np.arange(1, N + 1, dtype=np.int64)
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 5.0:
            clib.history_rows_naive(RULE, 0.5, h, Gh[(f + reps) % B], rows, n_threads=cores)
            reps += 1
        dt = time.perf_counter() - t0
        out['cpu_baseline'] = {'value': reps * len(rows) / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                               'sample': 'oracle/frx_oracle.c (C port of the reference algorithm, stencil rebuilt per row; the Python '
                                         'reference is ~80x slower per row, see the headline cpu_baseline), OpenMP, %d whole fields of '
                                         'N=16384 rows, %.2f s' % (reps, dt)}
    del G, Y
    torch.cuda.empty_cache()
    return out


def sub_cfg4(env, n, steps=20):
    """BASELINE cfg 4: assemble + batch-solve 1e5 systems per GPU over a 400 alpha x 250 length-scale sweep.  Metric: solves/s.
    (20 timed steps: every call starts with a host pass over the 1e5 settings, ~0.9 ms, that overlaps the previous call's
    kernels; over 5 steps the first call's exposed pass alone was 15 % of the n = 64 figure.)"""
    import numpy as np
    import torch
    import torch.distributed as dist
    from fractulus_b200 import batched, sweep
    world, rank, dev, ctx = env['world'], env['rank'], env['dev'], env['ctx']
    n_alpha, n_ls = 400, 250
    n_sys = n_alpha * n_ls
    h = 1.0 / (n + 1)
    alpha = np.repeat(np.linspace(0.05, 0.95, n_alpha), n_ls)
    res = np.tile(1.0 + (np.arange(n_ls) % 30), n_alpha)
    xg = torch.arange(1, n + 1, dtype=torch.float64, device=dev) * h
    rhs = torch.sin(math.pi * xg).repeat(n_sys, 1).contiguous()
    for _ in range(3):
        x = batched.assemble_solve(RULE, alpha, h, res, rhs)
    env['barrier']()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    ev0.record()
    for _ in range(steps):
        x = batched.assemble_solve(RULE, alpha, h, res, rhs)
    ev1.record()
    env['barrier']()
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
    solve_ms_per_step = ktimes['solve'][0] / steps              # ALL band-width bins of one step (one launch per bin)
    # Algorithmic flops.  A system that the pivot-free factorisation solves (definite, or indefinite and accepted on its
    # residual) is charged the band LDL^T: per column nb(nb+1)/2 FMAs + nb multipliers, and 4 nb + 1 flops for the two
    # substitutions (refinement steps are overhead, not algorithmic work); one that needs the partially pivoted band LU:
    # nb multipliers + nb x (2nb + 1) updates (x2) + the substitutions over the filled-in band.  Which systems are which is
    # the kernel's own count, per band-width bin (frx_solve_stats; checked against eigenvalues in tests/test_gpu_solve.py).
    nb = np.minimum(res + 1, n - 1)
    flop_lu = n * (2.0 * nb * (2 * nb + 1) + 3 * nb) + n * (4 * nb + 1)
    flop_ldl = n * (nb * (nb + 1.0) + nb) + n * (4 * nb + 1)
    n_stat, n_piv, piv_bins, refined_bins = ctx.solve_stats(per_bin=True)
    bins = np.digitize(res + 1, [7.5, 15.5, 23.5])
    flop = 0.0
    for b in range(4):
        m = bins == b
        if not m.any():
            continue
        share = 1.0 if piv_bins[b] < 0 else piv_bins[b] / float(m.sum())
        flop += share * float(flop_lu[m].sum()) + (1.0 - share) * float(flop_ldl[m].sum())
    dense_flop = n_sys * (2.0 / 3.0 * n ** 3 + 2.0 * n * n)
    achieved = flop / (solve_ms_per_step * 1e-3) / 1e12
    out = {'workload': 'cfg4: assemble + batch-solve %d systems per GPU, n=%d unknowns, 400 alpha in [0.05,0.95] x 250 length scales '
                       "(resolution 1..30), rule 'caputo' Riesz-Caputo operator, rhs sin(pi x); step = host pass over the settings + K5 stencil "
                       'weights + K4: in-kernel assembly, pivot-free LDL^T for every system (one thread per system at band width <= 7, '
                       'block LDL^T on DMMA above; indefinite systems kept only on their residual after refinement), partially pivoted '
                       'blocked band LU for what that rejects; the four band-width bins run as concurrent chains; PARITY UNPINNED (the '
                       'reference has no solver)' % (n_sys, n),
           'metric': 'fractional_bvp_solves_per_sec', 'value': world * n_sys / (ms_per_step * 1e-3), 'unit': 'solves/s',
           'ms_per_step': ms_per_step, 'steps': steps, 'gpu_launches': launches,
           'indefinite_systems': {'accepted_on_residual_after_refinement': int(sum(refined_bins)), 'solved_by_pivoted_lu': n_piv,
                                  'of': n_stat, 'refined_per_band_width_bin': refined_bins, 'pivoted_per_band_width_bin': piv_bins,
                                  'note': 'a pivot of the wrong sign in the pivot-free block LDL^T marks a system indefinite: it is kept '
                                          'only if ||b - A x|| <= 16 eps (max|a| ||x|| + ||b||) after <= 6 refinement steps, else the '
                                          'partially pivoted band LU solves it again'},
           'roofline': {'kernel': 'frx_k4_band_scalar (band width <= 7) + frx_k4_band_bldl (three wider bins and the narrow bin\'s indefinite systems) + frx_k4_band_dmma (fail lists)',
                        'bound': 'fp64', 'achieved': achieved,
                        'peak': env['peak_dmma'], 'unit': 'TFLOP/s', 'frac': achieved / env['peak_dmma'], 'traffic': None,
                        'kernel_ms': solve_ms_per_step, 'kernel_ms_is': 'ONE step, fork to join of the four concurrent bin chains (CUDA events on the context\'s stream)',
                        'algorithmic_flop_per_launch': flop, 'flop_if_every_system_took_the_pivoted_lu': float(flop_lu.sum()),
                        'dense_lu_flop_for_context': dense_flop,
                        'per_step_kernel_ms': {kk: v[0] / steps for kk, v in ktimes.items() if v[1]},
                        'note': 'flop = band LDL^T for the definite systems + pivoted band LU for the indefinite ones (not 2/3 n^3); the '
                                'kernels execute more than that (explicit 8x8 block inverses, full tiles at the band edge); peak = the '
                                'DMMA probe (DFMA probe %.2f TFLOP/s)' % env['peak_dfma']}}
    # residual of a sample through the kernel's own dense assembly
    idx = np.arange(0, n_sys, 997)
    A = batched.assemble_dense(RULE, alpha[idx], h, res[idx], n)
    r = torch.bmm(A, x[idx].unsqueeze(-1)).squeeze(-1) - rhs[idx]
    scale = (A.abs().amax(dim=(1, 2)) * x[idx].abs().amax(dim=1)).unsqueeze(-1)
    out['parity_spot_check'] = {'max_rel_residual_of_%d_sampled_systems' % len(idx): float((r.abs() / scale).max().item())}
    # e2e: host rhs in, host x out through the sweep API (pageable NumPy in, shared pinned host result out)
    rhs_h = batched.pinned_empty((n_sys, n))    # page-locked host right-hand sides
    torch.as_tensor(rhs_h).copy_(rhs)
    torch.cuda.synchronize()
    x_pinned = batched.pinned_empty((n_sys, n))
    if world == 1:
        def e2e_step():
            return sweep.assemble_solve_sweep(RULE, alpha, h, res, rhs_h)
    else:
        def e2e_step():
            return batched.assemble_solve_host(RULE, alpha, h, res, rhs_h, out=x_pinned)
    e2e_step()
    env['barrier']()
    t0 = time.perf_counter()
    for _ in range(3):
        xh = e2e_step()
    env['barrier']()
    e2e_s = (time.perf_counter() - t0) / 3
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    out['e2e'] = {'value': world * n_sys / e2e_s, 'unit': 'solves/s', 'ms_per_step': e2e_s * 1e3,
                  'h2d_bytes_per_step': (8 * n_sys * n + 24 * n_sys) * world, 'd2h_bytes_per_step': 8 * n_sys * n * world,
                  'api': 'sweep.assemble_solve_sweep (pinned NumPy rhs in, shared page-locked host result out, both used IN PLACE by the '
                         'kernels: every right-hand side is read once and every solution written once over PCIe while the batch solves)'
                         if world == 1 else
                         'batched.assemble_solve_host on every rank (its own systems; pinned NumPy in and out, used in place by the kernels)'}
    if rank == 0:
        from oracle import assemble
        k = 6
        pick = np.linspace(0, n_sys - 1, k).astype(int)
        t0 = time.perf_counter()
        worst = 0.0
        for s in pick:
            want, Amat = assemble.solve(RULE, alpha[s], h, int(res[s]), n, rhs_h[s])
            worst = max(worst, float(np.abs(np.asarray(xh)[s] - want).max() / np.abs(want).max() /
                                     (1e-12 + 20 * np.finfo(float).eps * np.linalg.cond(Amat))))
        dt = time.perf_counter() - t0
        out['parity_spot_check']['max_err_over_tolerance_vs_oracle_%d_systems' % k] = worst
        out['cpu_baseline'] = {'value': k / dt, 'unit': 'solves/s', 'cores': 1, 'kind': 'port',
                               'sample': 'oracle/assemble.py: reference stencils composed through the fdm shim into a dense matrix + '
                                         'numpy.linalg.solve (LAPACK), %d systems, %.2f s (no reference solver exists: parity unpinned)' % (k, dt)}
    del rhs, x
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fractulus_b200 import _lib, sweep

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- fractulus_b200 has no CPU path')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    ctx = _lib.context(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    env = {'world': world, 'rank': rank, 'local_rank': local_rank, 'dev': dev, 'ctx': ctx, 'barrier': barrier, 'args': args,
           'flush': torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)}
    A = args.alphas_per_gpu if args.alphas_per_gpu > 0 else (1 if world == 1 else CFG5_ALPHAS_PER_GPU)
    steps, warmup = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    head = bench_history(env, A, steps, warmup)
    clocks = sampler.window(*head['t_load']) if rank == 0 else None
    peak_dfma, peak_dmma = ctx.fp64_peak()
    env['peak_dfma'], env['peak_dmma'] = peak_dfma, peak_dmma

    configs = {}
    if not args.skip_configs:
        def guarded(name, fn, *a, **k):
            t0 = time.time()
            try:
                rec = fn(*a, **k)
            except Exception as ex:      # noqa: BLE001 -- a failing sub-record must not cost the headline line
                import traceback
                rec = {'error': repr(ex), 'traceback': traceback.format_exc()[-1500:]}
            if rank == 0:
                if isinstance(rec, dict) and 'error' not in rec:
                    rec['clocks'] = sampler.window(t0, time.time())
                configs[name] = rec
            barrier()
        if world == 1:
            guarded('cfg1', sub_cfg1, env)
        guarded('cfg3', sub_cfg3, env)
        for n in (64, 128, 256):
            guarded('cfg4_n%d' % n, sub_cfg4, env, n)
        if world == 1:
            def slice_rec():
                r = bench_history(env, CFG5_ALPHAS_PER_GPU, steps=3, warmup=3, warm_seconds=0.0)
                rec = public(r)
                rec['workload'] = ('cfg5 slice: %d alpha x N=1e5 on ONE GPU (the per-GPU share of the 1e4-alpha sweep on 8 GPUs); '
                                   'bench.py --gpus 8 runs cfg 5 itself' % CFG5_ALPHAS_PER_GPU)
                rec['unit'] = UNIT
                rec['roofline'] = history_roofline(r, CFG5_ALPHAS_PER_GPU, peak_dmma, peak_dfma, 3)
                return rec
            guarded('cfg5_slice', slice_rec)
            guarded('small_calls', sub_small_calls, env)
    if rank == 0:
        sampler.stop(0, 0)
        roofline = history_roofline(head, A, peak_dmma, peak_dfma, steps)
        if args.skip_cpu:
            cpu = {'skipped': '--skip-cpu (tuning run, not a bench line)'}
        else:
            arm = ReferenceArm()
            cpu, _ = arm.step(10.0)
            arm.close()
            try:
                cpu['c_port_for_context'] = c_port_rate()
            except Exception as ex:      # noqa: BLE001
                cpu['c_port_for_context'] = {'error': repr(ex)}
        line = {
            'metric': METRIC, 'value': head['value'], 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
            'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(world, A, gather=head['gather']),
            'e2e': head['e2e'], 'gpu_launches': head['gpu_launches'], 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
            'wall_s_timed_region': head['wall_s_timed_region'], 'ms_per_step_min': head['ms_per_step_min'],
            'ms_per_step_max': head['ms_per_step_max'], 'parity_spot_check': head.get('parity_spot_check'),
            'configs': configs,
        }
        if 'same_work_no_gather' in head:
            line['same_work_no_gather'] = head['same_work_no_gather']
        if world > 1:
            line['note'] = ('N>1 runs the cfg-5 slice (%d alpha per GPU per step); the N=1 headline is cfg 2 itself (one alpha per step, '
                            'a 0.3 ms launch that reaches 0.90 of the DMMA peak where the sweep reaches 0.96), so value(N)/(N*value(1)) '
                            'exceeds the true weak-scaling efficiency, which is same_work_no_gather.weak_scaling_efficiency_device'
                            % A) if A != 1 else None
        emit(line)
    if world > 1:
        sweep.release_plans()
        dist.barrier()
        dist.destroy_process_group()
    return 0


_JSON_LINES = []


def emit(line):
    _JSON_LINES.append(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=('b200', 'reference'))
    ap.add_argument('--alphas-per-gpu', type=int, default=0,
                    help='default: 1 at --gpus 1 (cfg 2), %d at --gpus > 1 (the cfg-5 sweep slice)' % CFG5_ALPHAS_PER_GPU)
    ap.add_argument('--nccl-gather', action='store_true', help='N > 1: gather with NCCL all_gather instead of the fused path')
    ap.add_argument('--skip-cpu', action='store_true', help='tuning only: skip the cpu_baseline leg')
    ap.add_argument('--skip-configs', action='store_true', help='tuning only: skip the sub-records of the other BASELINE configs')
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
