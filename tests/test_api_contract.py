"""CPU tests of the plugin API contract the package keeps from the reference (no device needed): the Application
constructor forms and the required-attribute check (tests/core/test_application.py), simple_setup_problem
(tests/core/test_simple_setup_problem.py:59-105), the argument validation of Mgrit (tests/core/test_mgrit.py:220-233,
core/mgrit.py:79-128 -- all raised before anything touches the device), the grid-transfer classes
(tests/core/test_grid_transfer_copy.py) and what a DeviceVector does with host data only."""
import copy

import numpy as np
import pytest

import cases as C
import pymgrit_b200 as P


class App(P.Application):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.vector_template = 2
        self.vector_t_start = 2

    def step(self, u_start, t_start, t_stop):
        return 1


class AppWithoutVectors(P.Application):
    def step(self, u_start, t_start, t_stop):
        pass


def test_application_constructor_forms():            # tests/core/test_application.py:37-100
    want = np.array([0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1])
    np.testing.assert_almost_equal(App(t_start=0, t_stop=1, nt=11).t, want)
    np.testing.assert_almost_equal(App(t_interval=np.linspace(0, 1, 11)).t, want)
    np.testing.assert_almost_equal(App(t_interval=np.linspace(0, 1, 11), t_start=0).t, want)
    a = App(t_interval=np.linspace(0, 1, 11))
    assert (a.t_start, a.t_end, a.nt) == (0.0, 1.0, 11)
    for bad in (dict(), dict(t_start=2, t_stop=5), dict(nt=4, t_stop=5), dict(nt=4, t_start=5), dict(t_interval=5)):
        with pytest.raises(Exception):
            App(**bad)


def test_required_attributes_are_enforced():         # core/application.py:17-29
    with pytest.raises(ValueError):
        AppWithoutVectors(t_start=0, t_stop=1, nt=11)


def test_simple_setup_problem():                     # tests/core/test_simple_setup_problem.py:59-105
    problem = P.simple_setup_problem(problem=App(t_start=0, t_stop=1, nt=101), level=3, coarsening=2)
    assert all(isinstance(p, App) for p in problem)
    for p, m in zip(problem, (1, 2, 4)):
        np.testing.assert_equal(p.t, np.linspace(0, 1, 101)[::m])
        assert (p.nt, p.t_start, p.t_end) == (len(p.t), 0, 1)
    small = P.simple_setup_problem(problem=App(t_start=0, t_stop=1, nt=2), level=2, coarsening=2)
    assert (small[1].nt, small[1].t_start, small[1].t_end) == (1, 0, 0)


def _heat(nt=65, **kw):
    args = dict(x_start=0, x_end=2, nx=5, a=1, rhs=C.heat_rhs, init_cond=C.heat_init, t_start=0, t_stop=2, nt=nt)
    args.update(kw)
    return P.Heat1D(**args)


@pytest.mark.parametrize('bad', [
    dict(cycle_type='Z'),                                               # tests/core/test_mgrit.py:220-223
    dict(t_norm=4),                                                     # :225-228
    dict(conv_crit=7),                                                  # :230-233 (mgrit.py:104-108)
    dict(output_lvl=3),                                                 # mgrit.py:89-90
    dict(cf_iter='1'),                                                  # mgrit.py:116-120
    dict(transfer=[P.GridTransferCopy(), P.GridTransferCopy()]),        # mgrit.py:79-80: one transfer per level pair
])
def test_mgrit_argument_validation(bad):
    with pytest.raises(Exception) as err:
        P.Mgrit(problem=[_heat()], **bad)
    assert 'CUDA' not in str(err.value)              # the argument check fired, not the missing device


def test_mgrit_level_validation():
    with pytest.raises(Exception):                   # coarse grid with more points than the fine one, mgrit.py:82-85
        P.Mgrit(problem=[_heat(17), _heat(65)])
    with pytest.raises(Exception):                   # coarse points that are not fine points, mgrit.py:86-88 / 212-214
        P.Mgrit(problem=[_heat(65), _heat(t_start=0.01, t_stop=1.99, nt=17)])
    with pytest.raises(Exception):                   # too few cf_iter entries, mgrit.py:110-115
        P.Mgrit(problem=[_heat(65), _heat(17), _heat(5)], cf_iter=[1])
    with pytest.raises(Exception):                   # an application without device kernels: no per-point fallback
        P.Mgrit(problem=[App(t_start=0, t_stop=1, nt=11)])
    with pytest.raises(Exception):                   # AT-MGRIT takes the global criteria only, at_mgrit.py:31-33
        P.AtMgrit(problem=[_heat(65), _heat(17)], k=2, conv_crit=2)


def test_grid_transfer_copy_and_abstract_base():     # tests/core/test_grid_transfer_copy.py, core/grid_transfer.py:15-55
    with pytest.raises(TypeError):
        P.GridTransfer()
    v = P.VectorHeat1D(3)
    v.set_values(np.array([1.0, 2.0, 3.0]))
    tr = P.GridTransferCopy()
    r, i = tr.restriction(v), tr.interpolation(v)
    np.testing.assert_array_equal(r.get_values(), [1, 2, 3])
    np.testing.assert_array_equal(i.get_values(), [1, 2, 3])
    r.set_values(np.zeros(3))
    np.testing.assert_array_equal(v.get_values(), [1, 2, 3])       # a clone, not an alias
    assert isinstance(P.GridTransferHeat1D(), P.DeviceGridTransfer)


def test_device_vector_host_side():
    """Values given on the host stay on the host until a kernel needs them: set/get, clone family, pack/unpack,
    deepcopy (core/vector.py:68-110) work without a device."""
    v = P.VectorHeat1D(4)
    assert not np.any(v.get_values()) and v.size == 4
    v.set_values(np.arange(4.0))
    c = v.clone()
    c.set_values(np.ones(4))
    np.testing.assert_array_equal(v.get_values(), np.arange(4.0))
    assert not np.any(v.clone_zero().get_values())
    assert v.clone_rand().get_values().shape == (4,)
    u = v.clone_zero()
    u.unpack(v.pack())
    np.testing.assert_array_equal(u.get_values(), np.arange(4.0))
    np.testing.assert_array_equal(copy.deepcopy(v).get_values(), np.arange(4.0))
    w = P.VectorHeat2D(3, 2)
    assert w.get_values().shape == (3, 2)
    pair = P.VectorHeat1D2Pts(3, 0.1)
    pair.set_values(np.ones(3), 2 * np.ones(3), 0.1)
    first, second, dtau = pair.get_values()
    assert first.tolist() == [1, 1, 1] and second.tolist() == [2, 2, 2] and dtau == 0.1 and pair.size == 3
    np.testing.assert_array_equal(pair.pack(), [[1, 1, 1], [2, 2, 2]])


def test_product_never_touches_the_oracle_or_the_reference():
    """oracle/ is test infrastructure and /root/reference is not on the GPU box: no file of the package (or of the C ABI)
    may import, open or link either; only tests/, bench.py's CPU legs and __graft_entry__.smoke() use the oracle."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for base in ('pymgrit_b200', 'include'):
        for dirpath, _, files in os.walk(os.path.join(root, base)):
            if os.sep + 'build' in dirpath or os.sep + 'lib' in dirpath:
                continue
            for f in files:
                if not f.endswith(('.py', '.cu', '.cuh', '.h')):
                    continue
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle\b', text, re.M) or 'mgrit_oracle' in text or \
                        re.search(r'open\([^)]*/root/reference', text) or 'sys.path' in text and '/root/reference' in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    bench = open(os.path.join(root, 'bench.py')).read()
    # nothing run on the GPU box reads the reference tree: the CPU arm drives the unmodified reference only where the tree
    # exists (the build container), behind reference_src(); the GPU arm never mentions it
    import inspect
    import bench as B
    assert '/root/reference' not in inspect.getsource(B.gpu_arm) + inspect.getsource(B.parity_check)
    assert 'os.path.isdir' in inspect.getsource(B.reference_src)
    entry = open(os.path.join(root, '__graft_entry__.py')).read()
    assert '/root/reference' not in entry


def test_every_case_reaches_the_device_boundary():
    """Without a GPU every parity case must get through the whole host-side setup that precedes the first device call
    (application constructors, right-hand-side analysis, argument checks, transfers) and then fail for the one reason
    that there is no CUDA device -- never with an AttributeError or a silent CPU result."""
    import logging
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA device present')
    from b200_util import b200_problem, b200_transfer
    for name, case in C.CASES.items():
        kw = dict(case['solver'])
        if 'transfer' in case:
            kw['transfer'] = b200_transfer(case)
        with pytest.raises(Exception) as err:
            if 'at_k' in case:
                P.AtMgrit(problem=b200_problem(case), k=case['at_k'], logging_lvl=logging.WARNING, **kw)
            else:
                P.Mgrit(problem=b200_problem(case), logging_lvl=logging.WARNING, **kw)
        assert 'CUDA' in str(err.value), (name, str(err.value))


@pytest.mark.parametrize('name', ['example_dahlquist', 'example_heat_1d', 'example_heat_1d_bdf2',
                                  'example_spatial_coarsening', 'example_heat_2d', 'example_at_mgrit',
                                  'example_allen_cahn'])
def test_examples_reach_the_device_boundary(name):
    """examples/*.py (the reference's examples with the import swapped): the host-side setup of each runs on the CPU and
    the solver constructor then stops for the one reason that there is no CUDA device."""
    import importlib
    import logging
    import os
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA device present')
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples')
    sys.path.insert(0, here)
    try:
        mod = importlib.import_module(name)
        kw = mod.build()
        cls = P.AtMgrit if 'k' in kw else P.Mgrit
        with pytest.raises(Exception) as err:
            cls(logging_lvl=logging.WARNING, **kw)
        assert 'CUDA' in str(err.value), str(err.value)
    finally:
        sys.path.remove(here)


def test_predicted_stop_of_the_queued_ahead_loop():
    """core/mgrit.py predicts_convergence: with two residuals known, the host waits for the residual of the cycle it has
    just queued instead of queueing another one, if linear convergence at the observed rate meets the tolerance."""
    from pymgrit_b200.core.mgrit import predicts_convergence
    conv = np.array([0.0, 3.9e-7, 3.5e-9, 0.0, 0.0])            # cfg 5: the third cycle is queued, two residuals are read
    assert predicts_convergence(conv, 2, 3, 1e-10)               # 3.5e-9 * 0.009 = 3.1e-11 < 1e-10
    assert not predicts_convergence(conv, 2, 3, 1e-12)
    assert not predicts_convergence(conv, 1, 2, 1e-10)           # one residual: no rate yet
    assert not predicts_convergence(conv, 2, 2, 1e-10)           # nothing queued beyond what was read
    assert not predicts_convergence(np.array([0.0, 1e-3, 2e-3, 0.0]), 2, 3, 1.0e-2)   # diverging: no prediction
    assert not predicts_convergence(np.array([0.0, 0.0, 1e-3, 0.0]), 2, 3, 1.0)       # no previous residual
    assert predicts_convergence(np.array([0.0, 1e-2, 1e-3, 0.0, 0.0]), 2, 4, 2e-5)    # two cycles queued ahead
