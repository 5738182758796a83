"""bench.py on a host without GPU: the reference arm (CPU oracle port) honours the driver's contract, ranks other
than 0 stay silent, and the product arm refuses to run without CUDA (there is no CPU fallback to time)."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, 'bench.py')


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.pop('RANK', None)
    e.pop('WORLD_SIZE', None)
    if env:
        e.update(env)
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run(['--impl', 'reference', '--gpus', '1', '--steps', '2', '--warmup', '1'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'graphs/s' and d['higher_is_better'] is True
    assert d['metric'] == 'protein-interface graphs/sec (GINet fwd+bwd, batch=64)'       # BASELINE.json's metric
    assert d['n_gpus'] == 1 and d['steps'] == 2 and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['dtype'] == 'f32' and d['data'] == 'synthetic'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] and d['cpu_baseline']['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'graphs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0
    # the same config keys the product arm prints (the driver compares the two arms' config)
    assert set(d['config']) == {'workload', 'path', 'batch_per_gpu', 'global_batch', 'step', 'l2', 'parallelism'}
    assert d['config']['batch_per_gpu'] == 64 and d['config']['workload'].startswith('cfg2')


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(['--impl', 'reference', '--gpus', '2', '--steps', '2', '--warmup', '1'], env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ''


def test_strong_scaling_flag_splits_the_global_batch():
    r = _run(['--impl', 'reference', '--workload', 'cfg2', '--gpus', '8', '--global-batch', '64', '--steps', '1', '--warmup', '1'])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['scaling'] == 'strong' and d['config']['batch_per_gpu'] == 8 and d['config']['global_batch'] == 64
    bad = _run(['--impl', 'reference', '--gpus', '8', '--global-batch', '60'])
    assert bad.returncode != 0 and 'multiple of --gpus' in bad.stderr
    both = _run(['--impl', 'reference', '--batch', '8', '--global-batch', '64'])
    assert both.returncode != 0


def test_product_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('this host has a GPU')
    r = _run(['--gpus', '1', '--steps', '2', '--warmup', '1'])
    assert r.returncode != 0
    assert 'no CPU fallback' in r.stderr


def test_step_bytes_follow_the_survey_figure():
    """roofline.achieved of the step kernel is SURVEY 8d's algorithmic figure (49.2 KB per cfg2 graph forward, x 2)
    times the graphs of a launch; the implementation bytes shrink by the fc1.weight rows with head v2."""
    import bench
    cfg = bench.workload_config('cfg2', None)
    _graphs, batches = bench.make_pool(cfg, 1, seed=0)
    b = batches[0]
    N, E = int(b.x.size(0)), int(b.edge_index.size(1))
    alg, impl = bench.step_bytes(b, N, E, cfg['feat'], 64, 10697)
    k0 = int(b.cluster1.numel())
    assert alg == 2 * (64 * (25600 + 16000 + 4000 + 1600 + 1600 + 4) + 8 * k0)
    assert abs(alg / 64 / 2 - 49204) < 200                      # ~49.2 KB per graph (K0 ~ 50)
    alg2, impl2 = bench.step_bytes(b, N, E, cfg['feat'], 64, 10697, 128 * 64 - 64)
    assert alg2 == alg and impl - impl2 == 2 * 4 * 64 * (128 * 64 - 64)


def test_step_kernel_traffic_reads_the_committed_capture():
    """roofline.traffic = dram__bytes_read + write per launch from the committed ncu capture, tagged with its file."""
    import bench
    v, src = bench.step_kernel_traffic('ginet_graph_step2_kernel', 'cfg2')
    assert src.startswith('profiles/') and os.path.exists(os.path.join(ROOT, src.split(' ')[0]))
    assert 2e6 < v < 2e7                                         # a few MB per launch: the compulsory input bytes
    v, src = bench.step_kernel_traffic('no_such_kernel', 'cfg2')
    assert v is None
