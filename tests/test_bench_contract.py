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


def test_class_fractions_use_the_launches_actually_profiled():
    """`tensor_core_classes` / `roofline.class_frac`: FLOPs of a kernel class = its PROFILED launches x one layer-operation each (a class that
    lost launches to another kernel must not keep their FLOPs), against both measured peaks."""
    sys.path.insert(0, ROOT)
    import bench
    peaks = {'tf_sustained': 1400.0, 'tf_burst': 1600.0}
    breakdown = {'conv_gemm_tc_kernel<256, 64, 4, 1>': (7.0, 560.0), 'conv_gemm_tc_kernel<128, 64, 3, 2>': (4.0, 520.0),
                 'conv_up4w_tc_kernel<2>': (3.0, 600.0), 'bn_act_bwd_apply_dense_kernel': (17.0, 970.0)}
    cls = bench.class_fractions(breakdown, 512, peaks)
    assert set(cls) == {'conv_gemm_tc_kernel', 'conv_up4w_tc_kernel'}
    g = cls['conv_gemm_tc_kernel']
    want = 11 * bench.LAYER_OP_FLOP * 512 / (1080.0 * 1e-6) / 1e12
    assert g['launches_per_step'] == 11 and abs(g['tflops'] - want) < 1e-9 * want
    assert abs(g['frac_of_sustained_peak'] - want / 1400.0) < 1e-12 and abs(g['frac_of_burst_peak'] - want / 1600.0) < 1e-12
    assert cls['conv_up4w_tc_kernel']['launches_per_step'] == 3
