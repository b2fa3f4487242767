"""The `--impl reference` arm of bench.py runs without a GPU (the driver launches it on the GPU
box, but nothing in it needs CUDA): one bounded step here, JSON contract checked."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_contract():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'leaves/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['steps'] == 1 and line['ms_per_step'] > 0
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and abs(cb['value'] - line['value']) < 1e-9
    assert line['e2e'] == dict(value=line['value'], unit='leaves/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert line['gpu_launches'] == 0 and line['metric'].startswith('simplices partitioned/sec')


def test_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1', CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=120, env=env,
                         cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ''
