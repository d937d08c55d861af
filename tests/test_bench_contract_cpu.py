"""
bench.py's reference arm (`--impl reference`: the CPU restatement of the reference loop, the only leg of bench.py that may
run without a GPU) prints ONE JSON line with the contract's keys; under a multi-rank launch only rank 0 prints.
"""

import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '1'], capture_output=True, text=True, timeout=600, env=env, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith('{')]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'forecast_steps_per_sec' and d['unit'] == 'forecast-steps/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['n_gpus'] == 1
    assert d['steps'] == 1 and d['warmup'] >= 1 and d['value'] > 0 and d['ms_per_step'] > 0
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] and 'sample' in d['cpu_baseline']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'}) == []
