"""CPU tests of bench.py: the reference arm end to end (it never touches the GPU) and the helpers of the GPU arm."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, MGRIT_BENCH_CPU_SAMPLE_NT='2049', MGRIT_BENCH_CPU_KIND='port')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1'],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'DOF/s' and d['higher_is_better'] is True and d['steps'] == 2
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['vs_baseline'] is None and d['dtype'] == 'f64'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'DOF/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'nt=1048577' in d['config']['workload'] and '2049' in d['cpu_baseline']['sample']
    assert set(d['config']) == {'workload'}                      # the same `config` as the GPU arm's workload: the sample is
    assert 0 < d['cpu_baseline']['window_fraction'] < 0.01       # described in cpu_baseline only
    assert d['cpu_baseline']['cpu_model'] and d['cpu_baseline']['iterations'] >= 1


def test_reference_arm_drives_the_unmodified_reference_when_it_is_there():
    """In the build container /root/reference/src exists: the arm then times pymgrit.Mgrit itself (kind 'reference');
    on the GPU box it does not exist and the oracle port is timed (kind 'port')."""
    import bench
    if bench.reference_src() is None:
        import pytest
        pytest.skip('no /root/reference on this box')
    env = dict(os.environ, MGRIT_BENCH_CPU_SAMPLE_NT='1025')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
    assert d['cpu_baseline']['kind'] == 'reference' and d['cpu_baseline']['cores'] == 1 and d['value'] > 0


def test_reference_arm_other_workloads():
    for wl, nt in (('cfg1', '101'), ('cfg4', '65'), ('cfg3', '17')):
        env = dict(os.environ, MGRIT_BENCH_CPU_SAMPLE_NT=nt, MGRIT_BENCH_CPU_KIND='port')
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                            '--workload', wl], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
        assert d['value'] > 0 and wl in d['config']['workload']


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', MGRIT_BENCH_CPU_SAMPLE_NT='2049')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_helpers():
    import bench
    import pymgrit_b200 as P
    nt, co = bench.workload_grid('cfg5')
    assert nt == 2 ** 20 + 1 and (nt - 1) % int(np.prod(co)) == 0
    levels = bench.hierarchy(P.Heat1D, 4097, (64, 16))
    assert [len(p.t) for p in levels] == [4097, 65, 5]
    assert np.array_equal(levels[1].t, levels[0].t[::64])
    assert 'coarsening 64x16' in bench.describe('cfg5', co)
    full = bench.ncu_traffic('f_relax', 16384, 64)
    assert full is not None and abs(full / (64 * 16384 * 8184) - 1) < 0.01          # DRAM bytes = algorithmic bytes
    assert bench.ncu_traffic('f_relax', 8192, 64) == full / 2
    assert bench.ncu_traffic('f_relax', 4096, 4) is None                             # captured for another interval length
    assert bench.ncu_traffic('no such sweep', 1, 64) is None
    s = bench.ClockSampler(0)
    s.proc = type('P', (), {'terminate': lambda self: None})()
    s.rows = [(1.0, ['1965', '1965', '700.0', 'Not Active', 'Not Active', 'Not Active', 'Active']),
              (2.0, ['1950', '1965', '900.0', 'Not Active', 'Not Active', 'Not Active', 'Not Active']),
              (9.0, ['1000', '1965', '100.0', 'Not Active', 'Not Active', 'Not Active', 'Not Active'])]
    s.mark_timed(0.5, 2.5)
    out = s.stop()
    assert out['sm_mhz'] == 1957.5 and out['sm_max_mhz'] == 1965.0 and out['reasons'] == ['sw_power_cap']
    assert out['samples'] == 2 and out['window'] == 'timed region'
