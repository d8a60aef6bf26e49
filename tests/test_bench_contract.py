"""CPU test of bench.py's contract pieces that need no GPU: the reference arm prints ONE JSON line with the keys the driver reads
(run at a tiny --scale so it takes seconds), rank != 0 prints nothing under torchrun, and roofline.traffic is only taken from an ncu
capture whose stamp carries the digest of the kernel sources it was compiled from."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(extra_env=None, *flags):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *flags], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.strip()]


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, 'oracle', '_ref')) and not os.path.exists(os.path.join(ROOT, 'oracle', 'oracle.py')),
                    reason='oracle not built')
def test_reference_arm_prints_one_contract_line_and_only_on_rank_zero():
    lines = run_bench(None, '--impl', 'reference', '--steps', '1', '--warmup', '0', '--scale', '0.002')
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['unit'] == 'edges/s' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['value'] > 0 and d['gpu_launches'] == 0 and d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert 'workload' in d['config'] and 'model' not in d['config']
    # under torchrun every rank but 0 exits without output
    assert run_bench({'RANK': '1', 'WORLD_SIZE': '2'}, '--impl', 'reference', '--steps', '1', '--warmup', '0', '--scale', '0.002') == []


def test_traffic_is_refused_without_a_matching_stamp(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    from gammagl_b200 import build
    tot, src = bench.ncu_traffic('c3', 'rmat')
    if tot is not None:                                   # the committed capture matches the kernel sources of this tree
        assert tot > 1e9 and src.startswith('profiles')
        assert json.load(open(os.path.join(ROOT, src) + '.stamp.json'))['kernel_digest'] == build.kernel_digest()
    assert bench.ncu_traffic('c2', 'rmat')[0] is None     # no capture for other workloads: null, never a guess
    monkeypatch.setattr(build, 'kernel_digest', lambda: 'something else')
    tot2, why = bench.ncu_traffic('c3', 'rmat')
    assert tot2 is None and ('stale' in why or 'no capture' in why or 'no stamp' in why)
