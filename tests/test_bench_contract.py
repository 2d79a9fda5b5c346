"""bench.py's JSON contract on CPU: the reference arm (`--impl reference`: the CPU oracle port of the reference optimiser, the only arm
that runs without a GPU) on the small plumbing workload, and the refusal of the product arm to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)


def test_reference_arm_line():
    r = _run('--impl', 'reference', '--workload', 'small', '--steps', '1', '--warmup', '1')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1                                                       # ONE JSON line
    j = json.loads(lines[0])
    assert j['impl'] == 'reference' and j['metric'] == 'person_frame_optimizer_iters_per_sec' and j['unit'] == 'person-frame-iters/s'
    assert j['higher_is_better'] is True and j['n_gpus'] == 1 and j['steps'] == 1 and j['warmup'] == 1 and j['vs_baseline'] is None
    assert j['value'] > 0 and abs(j['value'] - j['cpu_baseline']['value']) < 1e-9
    assert j['cpu_baseline']['kind'] == 'port' and j['cpu_baseline']['cores'] >= 1 and j['cpu_baseline']['sample']
    assert j['e2e'] == {'value': j['value'], 'unit': j['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert j['config']['workload'] and 'model' not in j['config']


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs and prints the reference arm."""
    r = _run('--impl', 'reference', '--workload', 'small', '--steps', '1', '--warmup', '1', '--gpus', '2', env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return                                                                   # on a GPU box the arm runs: covered by the bench itself
    r = _run('--workload', 'small', '--steps', '1', '--warmup', '3')
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
