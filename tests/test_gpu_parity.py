"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the reference-shaped Python API and
the C ABI, against (a) fixtures produced by the unmodified reference (tests/golden) and (b) the CPU oracle run live on the
same inputs.  Tolerances are SURVEY.md section 8c's: identical iteration count, solution within 1e-10 relative, residual
history within 1e-10 of the run's scale."""
import numpy as np
import pytest

import cases as C
from oracle_util import load_golden, run_oracle, assert_history_close, assert_solution_close
from b200_util import run_b200, solution_rows, b200_apps

pytestmark = pytest.mark.gpu

HEAVY = ('heat1d_cfg2',)
HAVE = None


def _cases(prefixes):
    return [k for k in C.CASES if k.startswith(prefixes) and k not in HEAVY]


def _check_against_golden(name):
    gold = load_golden(name)
    solver, info = run_b200(name)
    rows, norms = solution_rows(solver, gold['u_rows_idx'])
    scale = np.max(gold['u_norms']) * np.sqrt(len(gold['u_norms']))
    assert_history_close(info['conv'], gold['conv'], scale=scale)
    assert_solution_close(rows.reshape(gold['u_rows'].shape), norms, gold)
    return solver, info


@pytest.mark.parametrize('space', ['sine', 'node'])
@pytest.mark.parametrize('name', [k for k in _cases(('heat1d',)) if C.CASES[k]['app'] == 'heat1d' and 'at_k' not in C.CASES[k]])
def test_heat1d_against_reference_fixture(name, space, monkeypatch):
    """Every Heat1D fixture of the unmodified reference in both device representations of the level rows: sine
    coefficients (Phi diagonal, csrc/phi.cuh Heat1DSine; what Mgrit picks when it can) and node values (Toeplitz
    tridiagonal solve).  Spatial coarsening and non-separable right-hand sides stay in node space either way."""
    import pymgrit_b200 as P
    monkeypatch.setenv('MGB_HEAT1D_SINE', '1' if space == 'sine' else '0')
    solver, _ = _check_against_golden(name)
    case = C.CASES[name]
    sine_possible = 'transfer' not in case and solver.problem[0]._rhs_split.kind != 'dense'
    want = P._lib.APP_HEAT1D_SINE if (space == 'sine' and sine_possible) else P._lib.APP_HEAT1D
    assert all(p.kind == want for p in solver.problem)


@pytest.mark.parametrize('name', [k for k in C.CASES if C.CASES[k]['app'] == 'heat1d2pts'])
def test_heat1d_two_point_bdf_against_reference_fixture(name):
    """heat/heat_1d_2pts_bdf{1,2}.py: pair states, BDF2 over BDF1 levels (examples/example_heat_1d_bdf2.py)."""
    solver, _ = _check_against_golden(name)
    assert any(solver._fused_down) == (solver.weight_c == 1.0)


@pytest.mark.parametrize('name', [k for k in C.CASES if 'at_k' in C.CASES[k]])
def test_at_mgrit_against_reference_fixture(name):
    """core/at_mgrit.py on one time rank: local coarse grids on the coarsest level in one launch."""
    solver, info = _check_against_golden(name)
    if name == 'heat1d_atmgrit_test':              # tests/core/test_at_mgrit.py:33-45
        np.testing.assert_almost_equal(info['conv'], np.array([0.1767778, 0.01223507]))
    import pymgrit_b200 as P
    with pytest.raises(Exception):
        P.AtMgrit(problem=solver.problem, k=2, conv_crit=2)


@pytest.mark.parametrize('name', [k for k in C.CASES if C.CASES[k]['app'] == 'allencahn'])
def test_allen_cahn_imex_against_reference_fixture(name):
    """allen_cahn/allen_cahn.py (IMEX branch) on the batched path (core/batched.py, csrc/generic.cu): fixtures of the
    unmodified reference (examples/example_allen_cahn.py at nx = 32, an F-cycle on three levels, the jump criterion)."""
    solver, _ = _check_against_golden(name)
    assert solver._batched is not None and not any(solver._fused_down)


def test_spatial_coarsening_matches_reference_result_file():
    """examples/example_spatial_coarsening.py -> tests/mpi/results/spatial_coarsening (4 decimals there, tests/mpi/mpi.py:49)."""
    _, info = run_b200('heat1d_spatial_example')
    np.testing.assert_almost_equal(info['conv'], [0.033795341894154736, 0.0029793978719811257, 0.00032555028064712785,
                                                  4.042946916072736e-05, 4.93158057838271e-06, 6.178527940638919e-07,
                                                  7.708784717391436e-08], decimal=9)


def test_python_grid_transfer_is_rejected():
    """A transfer written in Python cannot run inside a sweep: the engine raises instead of falling back."""
    import pymgrit_b200 as P

    class Mine(P.GridTransfer):
        def restriction(self, u):
            return u.clone()

        def interpolation(self, u):
            return u.clone()

    prob = P.simple_setup_problem(P.Heat1D(nx=17, t_start=0, t_stop=1, nt=9, **C.HEAT), level=2, coarsening=2)
    with pytest.raises(Exception):
        P.Mgrit(problem=prob, transfer=[Mine()])
    with pytest.raises(Exception):      # sizes that do not nest
        P.Mgrit(problem=[P.Heat1D(nx=17, t_start=0, t_stop=1, nt=9, **C.HEAT),
                         P.Heat1D(nx=8, t_interval=np.linspace(0, 1, 5), **C.HEAT)], transfer=[P.GridTransferHeat1D()])


@pytest.mark.parametrize('name', _cases(('dahlquist', 'brusselator')))
def test_ode_against_reference_fixture(name):
    _check_against_golden(name)


@pytest.mark.parametrize('name', _cases(('advection',)))
def test_advection_against_reference_fixture(name):
    solver, _ = _check_against_golden(name)
    if name.endswith('_fourier'):       # these hierarchies end in a long level: solved in Fourier space (csrc/fourier.cu)
        assert type(solver._spectral.get(solver.lvl_max - 1)).__name__ == 'FourierSolve'


@pytest.mark.parametrize('name', _cases(('heat2d',)))
def test_heat2d_against_reference_fixture(name):
    if 'heat2d' not in b200_apps():
        pytest.skip('Heat2D device application not built yet')
    _check_against_golden(name)


def test_heat1d_cfg2_full_size():
    """BASELINE.json configs[1] in full: nx=1025, nt=16385, 3 levels, m=4, FCF V-cycle, tol 1e-10."""
    solver, info = _check_against_golden('heat1d_cfg2')
    assert len(info['conv']) == 3


@pytest.mark.parametrize('name', ['heat1d_small_v', 'heat1d_cfg2_nt1025', 'advection_example', 'heat1d_bdf1_small',
                                  'heat1d_bdf2_nonuniform'])
def test_against_live_oracle(name):
    """Same seeded inputs through the CPU oracle (C Thomas arithmetic) and the GPU, all level-0 points compared."""
    mg, ref = run_oracle(name, solver='c')
    solver, info = run_b200(name)
    rows, _ = solution_rows(solver)
    u_ref = mg.u[0]
    assert len(info['conv']) == len(ref['conv'])
    assert np.max(np.abs(rows - u_ref)) <= 1e-10 * np.max(np.abs(u_ref))
    scale = np.max(np.linalg.norm(u_ref.reshape(len(u_ref), -1), axis=1)) * np.sqrt(len(u_ref))
    assert_history_close(info['conv'], ref['conv'], scale=scale)
