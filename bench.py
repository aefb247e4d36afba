#!/usr/bin/env python
"""Benchmark of the S_n FFT hot path (contract: see the task brief / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload s8|s10|s12]

A step is one forward + one inverse transform of the whole synthetic batch.  Default workload is
BASELINE.json configs[1]: S_8, 1024 random functions (40 320 elements each, fp64), per GPU.
`value` counts element-transforms per second: each direction of each function contributes n!
elements, inputs resident in HBM.  `e2e` is the same through the public API from
pinned host buffers.  `sustained` is the same step looped for >= 2 s with its own clock record.
`coset_split` is north_star (4) at this N (also at N = 1): ONE S_12 function sharded by the coset
recursion over all ranks - parity against oracle-built columns on every rank, ms per step, the
per-kernel split, NVLink bytes and the exchange rate, and the prescribed NCCL reduce-scatter timed
alone for comparison.  Rank 0 prints ONE JSON line.
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

    def window(self, t0, t1):
        """Clock summary of the samples taken in [t0, t1] (call after stop())."""
        sm, power = [], []
        for ts, line in self.lines:
            if t0 <= ts <= t1 + 0.03:
                parts = [p.strip() for p in line.split(',')]
                try:
                    sm.append(float(parts[0]))
                    power.append(float(parts[2]))
                except (ValueError, IndexError):
                    continue
        return {'sm_mhz': statistics.median(sm) if sm else None, 'samples': len(sm),
                'power_w_max': max(power) if power else None}

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


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned
    host buffers of the e2e leg are allocated (first touch then places them on that node): with one
    process per GPU the host<->device copies of all ranks otherwise share one socket's memory and
    the inter-socket link.  Returns the node, or None when the topology is not exposed."""
    if os.environ.get('SNFFT_NO_NUMA_BIND'):
        return None
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(f'/sys/bus/pci/devices/{bus}/numa_node') as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f'/sys/devices/system/node/node{node}/cpulist') as fh:
            cpus = set()
            for part in fh.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:   # no GPU, no sysfs, no permission: placement is best effort
        return None


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
                   'n': n, 'batch_per_gpu': batch,
                   'l2': f'inputs {batch * math.factorial(n) * 8 / 1e6:.0f} MB per buffer, larger than the 126 MB L2'},
        'cpu_baseline': {'value': combined, 'unit': 'elements/s', 'cores': base.cores, 'kind': 'port',
                         'sample': base.sample_description(),
                         'forward_elements_per_s': statistics.mean(v[0] for v in vals),
                         'inverse_elements_per_s': statistics.mean(v[1] for v in vals)},
        'e2e': {'value': combined, 'unit': 'elements/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def _event_ms(fn, steps, barrier):
    """(ms for `steps` calls of fn, CUDA events on the current stream, barrier + sync both sides)."""
    import torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        fn()
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1)


def _max_over_ranks(value, world, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def run_coset_split(n, rank, world, local, steps, warmup, barrier, peak):
    """ONE function on S_n over all ranks, sharded by the coset recursion (north_star (4); reference
    pattern: cube_ft.py:82-106): parity on sparse-support inputs against oracle-built columns of
    rho_lambda(sigma) on every rank, then forward+inverse timed with the per-kernel split."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from snfft_b200 import get_plan, launch_count, timing_enable, timing_read
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import parity_tools as PT        # the oracle as the CHECKER (outside every timed region)

    dev = torch.device('cuda', local)
    nf = math.factorial(n)
    t0 = time.time()
    plan = get_plan(n, local)
    exchange = None
    if world > 1:
        from snfft_b200.sharded import ShardedPlan
        sp = ShardedPlan(n, rank, world, device=local)
        exchange = 'nccl all_to_all'
        if os.environ.get('SNFFT_SHARD_EXCHANGE', 'p2p') != 'nccl' and sp.enable_p2p():
            exchange = 'peer-memory push/pull kernels + device-side barriers (snfft_forward_sharded)'
        b, e = sp.local_elements()
    else:
        sp, b, e = None, 0, nf
    plan_seconds = time.time() - t0
    parts, dims, offs = plan.layout()

    spec_buf = torch.empty(nf if world == 1 else sp.sizes[rank][n], dtype=torch.float64, device=dev)
    back_buf = torch.empty(e - b, dtype=torch.float64, device=dev)
    work = torch.empty(nf, dtype=torch.float64, device=dev) if world == 1 else None

    def forward(x):
        if world == 1:
            return plan.forward(x, 'coset', out=spec_buf, work=work)
        return sp.forward_c(x, sp._p2p['wsk'], out=spec_buf) if getattr(sp, '_p2p', None) else sp.forward(x)

    def inverse(spec):
        if world == 1:
            return plan.inverse(spec, 'coset', out=back_buf, work=work)
        return sp.inverse_c(spec, sp._p2p['wsk'], out=back_buf) if getattr(sp, '_p2p', None) else sp.inverse(spec)

    # ---- parity: f = sum_j c_j delta_{sigma_j}  =>  f_hat(lambda)[:, col] = sum_j c_j rho_lambda(sigma_j)[:, col]
    perms, coefs = PT.sparse_support(n, count=4, seed=1000 + n)
    x = torch.zeros(e - b, dtype=torch.float64, device=dev)
    for p_, c_ in zip(perms, coefs):
        at = PT.O.coset_index(p_)
        if b <= at < e:
            x[at - b] += float(c_)
    spec = forward(x)
    irreps = PT.pick_irreps(parts, dims, big=3, small=3)
    if world == 1:
        def get_columns(i, cols):
            d = dims[i]
            return spec[offs[i]:offs[i] + d * d].reshape(d, d)[:, cols].cpu().numpy()
        owned = None
    else:
        views = sp.unpack(spec)

        def get_columns(i, cols):
            mine = sp.columns(i).tolist()
            return views[parts[i]][:, [mine.index(c) for c in cols]].cpu().numpy()

        def owned(i):
            return sp.columns(i)
    worst, checked = PT.check_columns(get_columns, parts, dims, perms, coefs, irreps, owned=owned)
    back = inverse(spec)
    rt_sparse = ((back - x).norm() / x.norm().clamp_min(1e-300)).item() if float(x.norm()) > 0 else 0.0
    worst = _max_over_ranks(worst, world, dev)
    del x, spec, back

    # ---- timing -------------------------------------------------------------------------------
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    x = torch.randn(e - b, dtype=torch.float64, device=dev, generator=gen)
    state = {}

    def step():
        state['spec'] = forward(x)
        state['back'] = inverse(state['spec'])

    for _ in range(warmup):
        step()
    barrier()
    err = ((state['back'] - x).norm() / x.norm()).item()
    launches0 = launch_count()
    timing_enable(True)
    timing_read()
    ms = _max_over_ranks(_event_ms(step, steps, barrier), world, dev) / steps
    kinds = timing_read()
    timing_enable(False)
    launches = launch_count() - launches0
    timed_out = bool(sp is not None and getattr(sp, '_p2p', None) and sp.barrier_timed_out())
    per_kind = {k: round(v[0] / steps, 4) for k, v in kinds.items() if v[1]}

    block = {
        'workload': f'S_{n} forward+inverse FFT of ONE function ({nf} elements, fp64), coset-split over {world} GPU(s)',
        'n': n, 'n_gpus': world, 'steps': steps, 'warmup': warmup, 'ms_per_step': ms,
        'value': 2 * nf / (ms * 1e-3), 'unit': 'elements/s', 'scaling': 'strong',
        'whole_step_frac_of_aggregate_hbm': (32.0 * nf / (ms * 1e-3) / 1e9) / (peak * world),
        'algorithmic_bytes_per_step': 32 * nf,
        'per_kernel_ms': per_kind,
        'kernel_ms_sum': round(sum(per_kind.values()), 4),
        'gpu_launches_per_step': launches / steps,
        'split_level_m': sp.m if sp else None,
        'exchange': exchange,
        'parity': {'check': 'sparse-support input (4 permutations): selected columns of f_hat(lambda) on every rank '
                            'against oracle-built columns of rho_lambda(sigma) (tests/parity_tools.py), global '
                            'column indices',
                   'max_rel_err': worst, 'tolerance': 1e-10, 'ok': bool(worst < 1e-10),
                   'irreps_rank0': [{'partition': list(p_), 'd': d_, 'columns': c_, 'rel_err': e_}
                                    for (p_, d_, c_, e_) in checked],
                   'round_trip_rel_err_random': err, 'round_trip_rel_err_sparse': rt_sparse,
                   'barrier_timed_out': timed_out},
        'plan_seconds': round(plan_seconds, 2),
    }
    if world > 1:
        nv = (world - 1) / world * 8.0 * nf / world
        ex_ms = per_kind.get('exchange', 0.0) / 2           # one push + one pull per step
        block['nvlink_bytes_per_rank_per_direction'] = nv
        block['exchange_ms_per_direction'] = ex_ms
        block['exchange_GBps_per_rank'] = nv / (ex_ms * 1e-3) / 1e9 if ex_ms > 0 else None
        # the collective the north star prescribes (cube_ft.py:102-106 as a reduce-scatter of
        # full-size partial spectra), timed alone for comparison with the one exchange above
        try:
            del state
            torch.cuda.empty_cache()
            full = torch.zeros(nf, dtype=torch.float64, device=dev)
            part = torch.empty(nf // world, dtype=torch.float64, device=dev)
            dist.reduce_scatter_tensor(part, full)
            rs = _max_over_ranks(_event_ms(lambda: dist.reduce_scatter_tensor(part, full), 2, barrier), world, dev) / 2
            block['nccl_reduce_scatter_full_spectrum_ms'] = rs
            block['nccl_reduce_scatter_note'] = ('reduce-scatter (sum, fp64) of a full-size partial spectrum per rank: '
                                                 'the collective alone of the prescribed variant, per direction')
            del full, part
        except Exception as exc:   # pragma: no cover
            block['nccl_reduce_scatter_full_spectrum_ms'] = None
            block['nccl_reduce_scatter_note'] = repr(exc)
    return block


def run_wreath(local, peak):
    """BASELINE configs[3]: the C3 wr S8 (2x2 cube) transform of a random value function at one of
    the top irreps the reference names in code, (2,3,3)/((1,1),(1,1,1),(2,1)), D = 1120
    (test_wreath.py:53-54).  Stage 1 = character sums over the orientations (C_3^7 DFT kernel,
    algorithmic bytes 8 |G| read + 16 n! N written), stage 2 = the sums over the Young subgroup per
    coset pair (fused gather kernel at this block size; gather + fp64 DMMA GEMM for large blocks)."""
    import torch
    from snfft_b200.wreath import WreathPlan
    dev = torch.device('cuda', local)
    alpha, parts = (2, 3, 3), ((1, 1), (1, 1, 1), (2, 1))
    t0 = time.time()
    wp = WreathPlan(alpha, parts, device=local)
    plan_seconds = time.time() - t0
    f = torch.randn((40320, 2187), dtype=torch.float64, device=dev,
                    generator=torch.Generator(device=dev).manual_seed(38))

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    ms_fwd, fh = timed(lambda: wp.forward(f), 5)
    ms_s1, _ = timed(lambda: wp.character_sums(f), 5)
    ms_inv, back = timed(lambda: wp.inverse(fh), 3)
    # Plancherel for this irrep alone cannot hold; the adjoint identity <f, inverse(F)> does:
    # inverse(forward(f)) is the projection of f on the irrep's isotypic component, so
    # <f, P f> = D ||f_hat||^2 / |G|
    lhs = (f * back.real).sum().item()
    rhs = wp.D * (fh.abs() ** 2).sum().item() / wp.order
    s1_bytes = 8.0 * wp.order + 16.0 * 40320 * wp.N
    flops = 2 * 2.0 * wp.N * wp.N * wp.nH * wp.block * wp.block
    return {'workload': 'C3 wr S8 forward / inverse at (2,3,3)/((1,1),(1,1,1),(2,1)), f on all 88 179 840 elements',
            'D': wp.D, 'cosets': wp.N, 'block': wp.block, 'young_subgroup': wp.nH,
            'forward_ms': ms_fwd, 'inverse_ms': ms_inv, 'elements_per_s_forward': wp.order / (ms_fwd * 1e-3),
            'stage1_ms': ms_s1, 'stage1_algorithmic_bytes': s1_bytes,
            'stage1_frac_of_hbm': s1_bytes / (ms_s1 * 1e-3) / 1e9 / peak,
            'stage2_ms': ms_fwd - ms_s1, 'stage2_gemm_flop': flops,
            'whole_forward_frac_of_hbm': (8.0 * wp.order + 16.0 * wp.D * wp.D) / (ms_fwd * 1e-3) / 1e9 / peak,
            'projection_identity_rel_err': abs(lhs - rhs) / abs(rhs), 'plan_seconds': round(plan_seconds, 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='s8', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-coset-split', action='store_true', help='skip the sharded S_12 block')
    ap.add_argument('--no-wreath', action='store_true', help='skip the C3 wr S8 block')
    ap.add_argument('--coset-split-n', type=int, default=12)
    ap.add_argument('--sustain-seconds', type=float, default=2.0)
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
    dev = torch.device('cuda', local)
    numa_node = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak, peak_src = measured_peak_gbs()
    n, batch = WORKLOADS[args.workload]
    nf = math.factorial(n)

    if args.workload != 's8':
        # one function sharded over all ranks is the whole line (strong scaling)
        blk = run_coset_split(n, rank, world, local, max(3, min(args.steps, 10)), args.warmup, barrier, peak)
        if rank == 0:
            line = {'metric': 'S_n FFT elements/sec', 'value': blk['value'], 'unit': 'elements/s', 'n_gpus': world,
                    'steps': blk['steps'], 'warmup': blk['warmup'], 'ms_per_step': blk['ms_per_step'],
                    'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
                    'data': 'synthetic', 'config': {'workload': blk['workload'], 'n': n},
                    'roofline': {'bound': 'hbm', 'achieved': blk['whole_step_frac_of_aggregate_hbm'] * peak * world,
                                 'peak': peak * world, 'unit': 'GB/s', 'frac': blk['whole_step_frac_of_aggregate_hbm'],
                                 'traffic': None, 'kernel': 'whole step (all kernels of both directions)',
                                 'peak_source': peak_src, 'per_kind_ms': blk['per_kernel_ms']},
                    'cpu_baseline': None, 'e2e': None, 'gpu_launches': int(blk['gpu_launches_per_step'] * blk['steps']),
                    'coset_split': blk}
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- primary: BASELINE configs[1], every rank transforms its own batch (replicas, weak scaling) ----
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    plan = get_plan(n, local)
    x = torch.randn(batch, nf, dtype=torch.float64, device=dev, generator=gen)
    spec = torch.empty_like(x)
    back = torch.empty_like(x)
    work = torch.empty_like(x)

    def step():
        plan.forward(x, 'coset', out=spec, work=work)
        plan.inverse(spec, 'coset', out=back, work=work)

    for _ in range(args.warmup):
        step()
    barrier()
    err = ((back - x).norm() / x.norm()).item()
    assert err < 1e-10, f'round trip failed: {err}'

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    launches0 = launch_count()
    timing_enable(True)
    timing_read()
    load_t0 = time.time()
    ms = _max_over_ranks(_event_ms(step, args.steps, barrier), world, dev)
    kinds = timing_read()
    timing_enable(False)
    launches = launch_count() - launches0
    elements_per_step = 2 * batch * nf * world
    value = elements_per_step * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel, timed live over the timed region --------------------
    dom = max(kinds, key=lambda k: kinds[k][0])
    dom_ms, dom_cnt = kinds[dom]
    bytes_per_launch = BYTES_PER_ELEMENT_PER_DIRECTION * batch * nf   # every pass reads + writes the batch
    achieved = bytes_per_launch / (dom_ms / max(1, dom_cnt) * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            traffic = json.load(fh).get(args.workload, {}).get(dom)
    except Exception:
        pass
    per_kind = {k: (v[0] / max(1, v[1]), v[1]) for k, v in kinds.items() if v[1]}
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': bytes_per_launch,
                'kernel_ms_per_launch': dom_ms / max(1, dom_cnt),
                'kernel_share_of_step': dom_ms / ms if ms > 0 else None,
                'per_kind_ms': {k: round(v[0] / max(1, args.steps), 4) for k, v in kinds.items() if v[1]},
                'per_kind_frac': {k: round(bytes_per_launch / (t * 1e-3) / 1e9 / peak, 4) for k, (t, c) in per_kind.items()},
                'whole_step_frac': (2 * bytes_per_launch / (ms / args.steps * 1e-3) / 1e9) / peak}

    # ---- the same step back to back for >= sustain_seconds: clocks under sustained load ----------
    sustained = None
    if args.sustain_seconds > 0:
        reps = max(args.steps, int(args.sustain_seconds * 1e3 / (ms / args.steps)) + 1)
        s_t0 = time.time()
        s_ms = _max_over_ranks(_event_ms(step, reps, barrier), world, dev)
        s_t1 = time.time()
        sustained = {'seconds': s_ms * 1e-3, 'steps': reps, 'ms_per_step': s_ms / reps,
                     'value': elements_per_step * reps / (s_ms * 1e-3), 't0': s_t0, 't1': s_t1}

    # ---- extras (ADVICE): forward only, and from / to lexicographic order ----------------------------
    f_ms = _max_over_ranks(_event_ms(lambda: plan.forward(x, 'coset', out=spec, work=work), args.steps, barrier), world, dev)

    def lex_step():
        plan.forward(x, 'lex', out=spec, work=work)
        plan.inverse(spec, 'lex', out=back, work=work)
    lex_step()
    l_ms = _max_over_ranks(_event_ms(lex_step, args.steps, barrier), world, dev)
    extras = {'forward_only_value': batch * nf * world * args.steps / (f_ms * 1e-3),
              'lex_order_value': elements_per_step * args.steps / (l_ms * 1e-3),
              'note': 'value counts forward + inverse in coset-major order; lex_order_value includes the two '
                      'reorder passes the reference enumeration order (perm2.sn) needs'}

    # ---- end to end from pinned host memory ------------------------------------------------------
    hx = torch.empty(tuple(x.shape), dtype=torch.float64).pin_memory()
    hx.copy_(x.cpu())
    hspec = torch.empty_like(hx).pin_memory()
    hback = torch.empty_like(hx).pin_memory()

    def e2e_step():
        # public host-buffer entry point: chunks pipelined over several streams so that H2D,
        # kernels and D2H overlap
        plan.roundtrip_host(hx, hspec, hback, order='coset',
                            chunks=int(os.environ.get('SNFFT_E2E_CHUNKS', 16)),
                            streams=int(os.environ.get('SNFFT_E2E_STREAMS', 4)))

    e2e_step()
    e2e_ms = _max_over_ranks(_event_ms(e2e_step, args.e2e_steps, barrier), world, dev)
    e2e_value = elements_per_step * args.e2e_steps / (e2e_ms * 1e-3)
    t_end = time.time()
    assert ((hback - hx).norm() / hx.norm()).item() < 1e-10
    clocks = sampler.stop(load_t0, t_end)   # device-timed, sustained and end-to-end regions
    if sustained is not None:
        sc = sampler.window(sustained.pop('t0'), sustained.pop('t1'))
        sustained['clocks'] = sc
    del hx, hspec, hback, x, spec, back, work
    torch.cuda.empty_cache()

    # ---- north_star (4): one S_12 function over all ranks, at every N including 1 ------------------
    coset_split = None
    if not args.no_coset_split:
        try:
            coset_split = run_coset_split(args.coset_split_n, rank, world, local, 5, 2, barrier, peak)
        except Exception as exc:     # the primary line must survive
            coset_split = {'error': repr(exc)}

    wreath = None
    if rank == 0 and not args.no_wreath:
        try:
            wreath = run_wreath(local, peak)
        except Exception as exc:
            wreath = {'error': repr(exc)}
    barrier()

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
               'inverse_elements_per_s': r_inv, 'seconds': secs,
               'note': 'the combined value includes replacing the reference direct per-element inverse '
                       '(tile/par_inv_ft.py) by a fast inverse; forward_elements_per_s vs '
                       'extras.forward_only_value is the algorithm-matched comparison'}

    if rank == 0:
        line = {
            'metric': 'S_n FFT elements/sec', 'value': value, 'unit': 'elements/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64',
            'data': 'synthetic',
            # the same dict as the reference arm prints (the driver compares them)
            'config': {'workload': f'S_{n} forward+inverse FFT, batch {batch} per GPU (configs[1])',
                       'n': n, 'batch_per_gpu': batch,
                       'l2': f'inputs {batch * nf * 8 / 1e6:.0f} MB per buffer, larger than the 126 MB L2'},
            'config_detail': {'elements_per_step': elements_per_step, 'directions_per_step': 2,
                              'order': 'coset-major',
                              'parallelism': f'batch-replicas x{world}' if world > 1 else 'single GPU',
                              'round_trip_rel_err': err, 'host_numa_node_rank0': numa_node},
            'roofline': roofline,
            'cpu_baseline': cpu,
            'e2e': {'value': e2e_value, 'unit': 'elements/s',
                    'h2d_bytes_per_step': batch * nf * 8, 'd2h_bytes_per_step': 2 * batch * nf * 8},
            'gpu_launches': launches,
            'clocks': clocks,
            'sustained': sustained,
            'extras': extras,
            'coset_split': coset_split,
            'wreath': wreath,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
