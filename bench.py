#!/usr/bin/env python
"""Benchmark of the S_n FFT hot path (contract: see the task brief / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload s8|s10|s12]

A step is one forward + one inverse transform of the whole synthetic batch.  Default workload is
BASELINE.json configs[1]: S_8, 1024 random functions (40 320 elements each, fp64), per GPU.
`value` counts element-transforms per second: each direction of each function contributes n!
elements, inputs resident in HBM.  `e2e` is the same through the public API from
pinned host buffers.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, functions per GPU)
    's8': (8, 1024),
    's10': (10, 8),
    's12': (12, 1),
}
BYTES_PER_ELEMENT_PER_DIRECTION = 16  # 8 B read + 8 B written


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                 '-lms', '20', '-i', str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        # samples taken while the GPU was under load (the timed regions), else everything we have
        inside = [ln for (ts, ln) in self.lines if t0 is not None and t0 <= ts <= t1 + 0.03]
        in_region = len(inside)
        for line in (inside if inside else [ln for _, ln in self.lines]):
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'samples_in_timed_region': in_region, 'reasons': sorted(reasons)}


def run_reference(args, rank, world):
    """The reference arm: the oracle port of the reference's CPU algorithms on all host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import cpu_baseline
    n, batch = WORKLOADS[args.workload]
    if n > 8:
        print(json.dumps({'impl': 'reference', 'unavailable':
                          f'reference algorithm infeasible at n={n} (tableau brute force, direct inverse)'}))
        return
    base = cpu_baseline.Baseline(n)
    for i in range(args.warmup):
        base.step(seed=100 + i)
    t0 = time.perf_counter()
    vals = [base.step(seed=i) for i in range(args.steps)]
    total = time.perf_counter() - t0
    base.close()
    combined = statistics.mean(v[2] for v in vals)
    line = {
        'impl': 'reference', 'metric': 'S_n FFT elements/sec', 'value': combined, 'unit': 'elements/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / max(1, args.steps), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'S_{n} forward+inverse FFT, batch {batch} per GPU (configs[1])',
                   'n': n, 'batch_per_gpu': batch},
        'cpu_baseline': {'value': combined, 'unit': 'elements/s', 'cores': base.cores, 'kind': 'port',
                         'sample': base.sample_description(),
                         'forward_elements_per_s': statistics.mean(v[0] for v in vals),
                         'inverse_elements_per_s': statistics.mean(v[1] for v in vals)},
        'e2e': {'value': combined, 'unit': 'elements/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='s8', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from snfft_b200 import get_plan, launch_count, timing_enable, timing_read

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n, batch = WORKLOADS[args.workload]
    nf = math.factorial(n)
    gen = torch.Generator(device=f'cuda:{local}').manual_seed(1234 + rank)
    # s10 / s12 on several GPUs: ONE function sharded by the coset recursion (strong scaling);
    # otherwise every rank transforms its own batch (replicas, weak scaling)
    sharded = world > 1 and args.workload != 's8'
    if sharded:
        from snfft_b200.sharded import ShardedPlan
        batch = 1
        plan = ShardedPlan(n, rank, world, device=local)
        # the exchange as one kernel per direction over NVLink peer memory unless SNFFT_SHARD_EXCHANGE=nccl
        exchange = 'nccl all_to_all'
        if os.environ.get('SNFFT_SHARD_EXCHANGE', 'p2p') != 'nccl' and plan.enable_p2p():
            exchange = 'peer-memory push/pull kernel (snfft_shard_exchange)'
        b, e = plan.local_elements()
        x = torch.randn(e - b, dtype=torch.float64, device=f'cuda:{local}', generator=gen)
        state = {}

        def step():
            state['spec'] = plan.forward(x)
            state['back'] = plan.inverse(state['spec'])
    else:
        plan = get_plan(n, local)
        x = torch.randn(batch, nf, dtype=torch.float64, device=f'cuda:{local}', generator=gen)
        spec = torch.empty_like(x)
        back = torch.empty_like(x)
        work = torch.empty_like(x)

        def step():
            plan.forward(x, 'coset', out=spec, work=work)
            plan.inverse(spec, 'coset', out=back, work=work)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    if sharded:
        back = state['back']
    err = ((back - x).norm() / x.norm()).item()
    assert err < 1e-10, f'round trip failed: {err}'

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    launches0 = launch_count()
    timing_enable(True)
    timing_read()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    load_t0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    kinds = timing_read()
    timing_enable(False)
    launches = launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=f'cuda:{local}')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    # both directions; replicas: every rank has its own batch, sharded: one function in total
    elements_per_step = 2 * nf if sharded else 2 * batch * nf * world
    value = elements_per_step * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel, timed live over the timed region --------------------
    peak, peak_src = measured_peak_gbs()
    dom = max(kinds, key=lambda k: kinds[k][0])
    dom_ms, dom_cnt = kinds[dom]
    # every pass reads + writes this rank's share of the data
    bytes_per_launch = BYTES_PER_ELEMENT_PER_DIRECTION * (nf // world if sharded else batch * nf)
    achieved = bytes_per_launch / (dom_ms / max(1, dom_cnt) * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            traffic = json.load(fh).get(args.workload, {}).get(dom)
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': bytes_per_launch,
                'kernel_ms_per_launch': dom_ms / max(1, dom_cnt),
                'kernel_share_of_step': dom_ms / ms if ms > 0 else None,
                'per_kind_ms': {k: round(v[0] / max(1, args.steps), 4) for k, v in kinds.items() if v[1]},
                'whole_step_frac': (2 * bytes_per_launch / (ms / args.steps * 1e-3) / 1e9) / peak}

    # ---- end to end from pinned host memory ------------------------------------------------------
    hx = torch.empty(tuple(x.shape), dtype=torch.float64).pin_memory()
    hx.copy_(x.cpu())
    if sharded:
        hspec = torch.empty(state['spec'].shape, dtype=torch.float64).pin_memory()
    else:
        hspec = torch.empty_like(hx).pin_memory()
    hback = torch.empty_like(hx).pin_memory()

    def e2e_step():
        if sharded:
            x.copy_(hx, non_blocking=True)
            sp = plan.forward(x)
            hspec.copy_(sp, non_blocking=True)
            hback.copy_(plan.inverse(sp), non_blocking=True)
        else:
            # public host-buffer entry point: chunks pipelined over several streams so that H2D,
            # kernels and D2H overlap
            plan.roundtrip_host(hx, hspec, hback, order='coset',
                                chunks=int(os.environ.get('SNFFT_E2E_CHUNKS', 16)),
                                streams=int(os.environ.get('SNFFT_E2E_STREAMS', 4)))

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f'cuda:{local}')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = elements_per_step * args.e2e_steps / (t.item() * 1e-3)
    clocks = sampler.stop(load_t0, time.time())   # device-timed and end-to-end regions
    assert ((hback - hx).norm() / hx.norm()).item() < 1e-10

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and n <= 8:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import cpu_baseline
        base = cpu_baseline.Baseline(n)
        base.step(seed=99)
        r_fwd, r_inv, combined, secs = base.step(seed=0)
        base.close()
        cpu = {'value': combined, 'unit': 'elements/s', 'cores': base.cores, 'kind': 'port',
               'sample': base.sample_description(), 'forward_elements_per_s': r_fwd,
               'inverse_elements_per_s': r_inv, 'seconds': secs}

    if rank == 0:
        line = {
            'metric': 'S_n FFT elements/sec', 'value': value, 'unit': 'elements/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if sharded else 'weak', 'vs_baseline': None,
            'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': f'S_{n} forward+inverse FFT, batch {batch} per GPU '
                                   f'({"configs[1]" if args.workload == "s8" else args.workload})',
                       'n': n, 'batch_per_gpu': batch, 'elements_per_step': elements_per_step,
                       'directions_per_step': 2, 'order': 'coset-major',
                       'l2': f'inputs {batch * nf * 8 / 1e6:.0f} MB per buffer, larger than the 126 MB L2'
                             if batch * nf * 8 > 126e6 else 'working set fits L2: not an HBM number',
                       'parallelism': (f'coset-split column shards x{world}, one exchange: {exchange}' if sharded else
                                       f'batch-replicas x{world}' if world > 1 else 'single GPU'),
                       'round_trip_rel_err': err},
            'roofline': roofline,
            'cpu_baseline': cpu,
            'e2e': {'value': e2e_value, 'unit': 'elements/s',
                    'h2d_bytes_per_step': x.numel() * 8, 'd2h_bytes_per_step': 2 * x.numel() * 8},
            'gpu_launches': launches,
            'clocks': clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
