"""The JSON contract of bench.py, checked without a GPU: the reference arm runs here (one bounded CPU step), and the line committed from the
final B200 run of the round must carry every key the driver and the judge read."""
import glob
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
             'e2e', 'gpu_launches', 'cpu_baseline'}


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
    assert BASE_KEYS <= set(line) and line['impl'] == 'reference' and line['metric'] == 'dcgan_train_images_per_sec' and line['unit'] == 'images/s'
    assert line['value'] > 0 and line['higher_is_better'] is True and line['vs_baseline'] is None and line['gpu_launches'] == 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1 and line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in line['config'] and 'model' not in line['config']


def test_committed_b200_line_carries_every_contract_key():
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r02_*_bench_1gpu_default*.json')))
    assert files, 'the final 1-GPU bench line of the round is kept under profiles/'
    line = json.loads([l for l in open(files[-1]) if l.startswith('{')][-1])
    assert BASE_KEYS | {'roofline', 'clocks', 'more_configs', 'top_kernels', 'tensor_core_classes'} <= set(line)
    assert line['n_gpus'] == 1 and line['dtype'] == 'bf16' and line['data'] == 'synthetic' and line['scaling'] == 'weak' and line['vs_baseline'] is None
    roof = line['roofline']
    assert roof['bound'] == 'tensor' and roof['unit'] == 'TFLOP/s' and abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-6 and roof['traffic'] is None
    assert set(line['e2e']) >= {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} and line['e2e']['h2d_bytes_per_step'] > 5e7
    assert line['e2e']['value'] != line['value'] and line['gpu_launches'] > 0
    assert set(line['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'} and not {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'} & set(line['clocks']['reasons'])
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['value'] > 0
    names = ' '.join(m.get('model', '') for m in line['more_configs'])
    assert 'WGAN-GP' in names and 'CGAN' in names and any(m.get('nc') == 3 and 'model' not in m for m in line['more_configs'])
