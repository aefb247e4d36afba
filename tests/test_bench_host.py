"""bench.py on a box without a GPU: the reference arm's JSON line (the driver runs
`bench.py --impl reference` and compares its `config` with our arm's) and the host-side helpers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['higher_is_better'] is True and line['dtype'] == 'f64'
    assert line['metric'] == 'S_n FFT elements/sec' and line['unit'] == 'elements/s' and line['value'] > 0
    assert line['steps'] == 1 and line['n_gpus'] == 1
    assert line['config']['n'] == 8 and line['config']['batch_per_gpu'] == 1024
    cpu = line['cpu_baseline']
    assert cpu['kind'] == 'port' and cpu['cores'] >= 1 and cpu['value'] == line['value'] and cpu['sample']
    e2e = line['e2e']
    assert e2e['value'] == line['value'] and e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0


def test_reference_arm_under_torchrun_only_rank0_works():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith('{')]


def test_host_helpers_without_a_gpu():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.bind_to_gpu_numa_node(0) is None or isinstance(bench.bind_to_gpu_numa_node(0), int)
    peak, source = bench.measured_peak_gbs()
    assert peak > 1000 and source
