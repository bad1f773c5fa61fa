"""Pins the CPU oracle (oracle/mgrit_oracle.py) against the reference: its known-answer vectors, its golden
residual files, and fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import cases as C
from oracle import mgrit_oracle as O
from oracle_util import run_oracle, load_golden, assert_history_close, assert_solution_close

FAST = [k for k, c in C.CASES.items() if not c.get('slow') and k not in ('advection_cfg4_small', 'heat2d_cfg3_small',
                                                                        'advection_nx4096_short')]


# ---- known-answer vectors from the reference's own unit tests --------------------------------
def test_heat1d_step_known_answer():      # tests/heat/test_heat_1d.py:31-42
    p = O.Heat1DOracle(a=1, init_cond=lambda x: 2 * x, x_start=0, x_end=1, nx=6, t_start=0, t_stop=1, nt=11)
    want = np.array([0.28164, 0.51593599, 0.63660638, 0.53191933])
    np.testing.assert_almost_equal(p.phi(p.u0, 0, 0.1), want)
    p.solver = 'c'
    np.testing.assert_almost_equal(p.phi(p.u0, 0, 0.1), want)


def test_heat2d_step_known_answer():      # tests/heat/test_heat_2d.py:230-249
    p = O.Heat2DOracle(a=1, x_start=0, x_end=1, y_start=3, y_end=4, nx=5, ny=5, rhs=lambda x, y, t: 2 * x * y,
                       t_start=0, t_stop=1, nt=11)
    want = np.array([[0., 0., 0., 0., 0.], [0., 0.06659024, 0.08719337, 0.07227713, 0.],
                     [0., 0.11922399, 0.15502696, 0.12990086, 0.], [0., 0.12666875, 0.16193148, 0.1391124, 0.],
                     [0., 0., 0., 0., 0.]])
    np.testing.assert_almost_equal(p.phi(p.u0, 0, 0.1), want)


def test_advection_step_known_answer():   # tests/advection/test_advection_1d.py:33-45
    p = O.Advection1DOracle(c=1, x_start=0, x_end=1, nx=6, t_start=0, t_stop=1, nt=11)
    want = np.array([0.868043, 0.92987396, 0.87805385, 0.75780217, 0.604129])
    np.testing.assert_almost_equal(p.phi(p.u0, 0, 0.1), want)
    p.solver = 'c'
    np.testing.assert_almost_equal(p.phi(p.u0, 0, 0.1), want)


def test_dahlquist_step_known_answer():   # tests/dahlquist/test_dahlquist.py:55-62
    p = O.DahlquistOracle(t_start=0, t_stop=1, nt=11)
    np.testing.assert_almost_equal(p.phi(p.u0, 0, 0.1), 0.9090909090909091)


def test_brusselator_step_known_answer():  # tests/brusselator/test_brusselator.py:24-31 (RK4 step from (0, 0))
    p = O.BrusselatorOracle(t_start=0, t_stop=1, nt=11)
    np.testing.assert_almost_equal(p.phi(np.zeros(2), 0, 0.1), np.array([0.08240173, 0.01319825]))


def test_allen_cahn_imex_step_known_answer():   # tests/allen_cahn/test_allen_cahn.py:123-146 (test_heat_2d_step_imex)
    prob = O.AllenCahnOracle(nx=3, eps=3, radius=0.25, method='IMEX', nu=1, t_start=0, t_stop=1, nt=11)
    np.testing.assert_almost_equal(prob.u0, np.array([[-0.10732614, -0.05885746, -0.10732614],
                                                       [-0.05885746, 0.05885746, -0.05885746],
                                                       [-0.10732614, -0.05885746, -0.10732614]]))
    np.testing.assert_almost_equal(prob.phi(prob.u0, 0, 0.1), np.array([[-0.07997795, -0.0640509, -0.07997795],
                                                                         [-0.0640509, -0.03719789, -0.0640509],
                                                                         [-0.07997795, -0.0640509, -0.07997795]]))
    np.testing.assert_equal(prob.space_disc.toarray()[0], [-36., 9., 9., 9., 0., 0., 9., 0., 0.])   # :41-49


def test_heat1d_2pts_step_known_answers():
    # tests/heat/test_heat_1d_2pts_bdf1.py:35-54
    p = O.Heat1D2PtsOracle(a=1, init_cond=lambda x: 2 * x, x_start=0, x_end=1, nx=11, dtau=0.1, method='BDF1', t_start=0,
                           t_stop=1, nt=11)
    got = p.phi(p.u0, 0, 0.1)
    np.testing.assert_almost_equal(got[0], np.array(
        [0.14498001, 0.28445802, 0.41238183, 0.52154382, 0.6028602, 0.6444626, 0.63051125, 0.53961104, 0.34267192]))
    np.testing.assert_almost_equal(got[1], np.array(
        [0.08691756, 0.16802887, 0.23749726, 0.2894772, 0.31825048, 0.31856279, 0.28628511, 0.21958482, 0.12088191]))
    # tests/heat/test_heat_1d_2pts_bdf2.py:37-62
    p = O.Heat1D2PtsOracle(a=1, init_cond=lambda x: 2 * x, x_start=0, x_end=1, nx=11, dtau=0.1, method='BDF2', t_start=0,
                           t_stop=1, nt=5)
    np.testing.assert_almost_equal(p.u0[0], np.array([0.2, 0.4, 0.6, 0.8, 1., 1.2, 1.4, 1.6, 1.8]))
    np.testing.assert_almost_equal(p.u0[1], np.array(
        [0.15656217, 0.30443677, 0.43319873, 0.52860043, 0.56972221, 0.52478844, 0.34481236, -0.04620125, -0.76645512]))
    got = p.phi(p.u0, 0, 0.2)
    np.testing.assert_almost_equal(got[0], np.array(
        [0.07115547, 0.13167183, 0.17105162, 0.1794494, 0.1490445, 0.07705183, -0.02834074, -0.1369469, -0.17685485]))
    np.testing.assert_almost_equal(got[1], np.array(
        [0.01235156, 0.02015287, 0.01986458, 0.01000559, -0.00781242, -0.02812508, -0.04182745, -0.03889518,
         -0.01671786]))


# ---- golden residual files of the reference (tests/mpi/results/*, compared at 4 decimals there) ----
REF_FILES = {
    'dahlquist_cfg1': [7.186185937025429e-05, 1.246106707585954e-06, 2.1015566149418615e-08, 3.1441273895579124e-10,
                       3.975216519949153e-12],
    'heat1d_example': [1.6743745803519487, 0.08233110058226714, 0.004141227358231137, 0.00020796732945572877,
                       1.0239713477386588e-05, 4.841251066261163e-07, 2.1337164720045163e-08],
    'heat1d_weighted': [1.379505913515878, 0.05379386198383165, 0.0021437634917656637, 8.410927592723663e-05,
                        3.135559182803784e-06, 1.0582423513330225e-07, 2.9689185659602585e-09],
    'dahlquist_ml1': [0.00019401363972256573, 7.975571640343785e-06, 2.99332402020636e-07, 8.881441953215148e-09,
                      1.9391939687035342e-10, 3.0368067573802767e-12],
    'dahlquist_procs_wo_pts': [0.007693147105384459, 0.0005069883607615411, 1.2469121704416458e-05,
                               1.7860071305544555e-17],
    'dahlquist_varying': [0.037311841611405, 0.003124171062320715, 3.129166834664884e-05, 1.8514542798812671e-07,
                          4.995916285724713e-10, 4.82164655680165e-13],
    'dahlquist_integrators': [0.0003079175789847042, 1.1038989829978764e-05, 3.8490474036390616e-07,
                              1.1905574823105624e-08],
    'brusselator_example': [0.014222592121653527, 8.204212669361892e-05, 1.1265397600214108e-07,
                            3.357814854491385e-10],
    'heat2d_example': [5.372296482469411e-15],
    'heat1d_spatial_example': [0.033795341894154736, 0.0029793978719811257, 0.00032555028064712785,
                               4.042946916072736e-05, 4.93158057838271e-06, 6.178527940638919e-07,
                               7.708784717391436e-08],
}


@pytest.mark.parametrize('name', sorted(REF_FILES))
def test_reference_result_files(name):
    _, info = run_oracle(name)
    np.testing.assert_almost_equal(info['conv'], REF_FILES[name], decimal=4)   # tests/mpi/mpi.py:49


def test_reference_core_test_values():    # tests/core/test_mgrit.py:59-70
    _, info = run_oracle('heat1d_testmgrit')
    np.testing.assert_almost_equal(info['conv'], np.array([0.00267692, 0.00018053]))


def test_one_level_is_time_stepping():    # tests/core/test_mgrit.py:72-84
    mg, info = run_oracle('heat1d_onelevel')
    assert len(info['conv']) == 0
    np.testing.assert_almost_equal(mg.u[0], O.time_stepping(mg.problem[0]))


# ---- fixtures generated from the unmodified reference in the build container ---------------------
@pytest.mark.parametrize('name', FAST)
def test_against_reference_fixture(name):
    gold = load_golden(name)
    mg, info = run_oracle(name)
    u = mg.u[0]
    norms = [np.linalg.norm(x) for x in u]
    assert_history_close(info['conv'], gold['conv'], scale=np.max(gold['u_norms']), rtol=1e-12)
    assert_solution_close(u[gold['u_rows_idx']], norms, gold, rtol=1e-12)


@pytest.mark.parametrize('name', ['heat1d_cfg2_nt1025', 'advection_example', 'heat1d_small_f_cf2'])
def test_c_phi_matches_superlu_solver(name):
    """Thomas / forward substitution in C against the SuperLU arithmetic (SURVEY.md 8c tolerance)."""
    gold = load_golden(name)
    mg, info = run_oracle(name, solver='c')
    norms = [np.linalg.norm(x) for x in mg.u[0]]
    assert_history_close(info['conv'], gold['conv'], scale=np.max(gold['u_norms']) * np.sqrt(len(norms)))
    assert_solution_close(mg.u[0][gold['u_rows_idx']], norms, gold)


def test_phi_step_fixtures():
    g = load_golden('phi_steps')
    h = O.Heat1DOracle(x_start=0, x_end=1, nx=1025, a=1, init_cond=C.heat_init, rhs=C.heat_rhs, t_start=0, t_stop=2,
                       nt=5)
    for solver, tol in (('spsolve', 1e-14), ('c', 1e-11)):
        h.solver = solver
        for k in range(5):
            dt = float(g[f'heat1d_1025/dt{k}'][0])
            got = h.phi(h.u0, 0.3, 0.3 + dt)
            ref = g[f'heat1d_1025/out{k}']
            assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref))
    a = O.Advection1DOracle(c=1, x_start=-1, x_end=1, nx=4096, t_start=0, t_stop=2, nt=5)
    for solver, tol in (('spsolve', 1e-14), ('c', 1e-12)):
        a.solver = solver
        for k in range(3):
            dt = float(g[f'advection_4096/dt{k}'][0])
            got = a.phi(a.u0, 0.0, dt)
            ref = g[f'advection_4096/out{k}']
            assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref))
    h2 = O.Heat2DOracle(x_start=0, x_end=1, y_start=0, y_end=1, nx=65, ny=49, a=1, rhs=C.heat2d_rhs,
                        init_cond=C.heat2d_init, bc_left=1.0, bc_top=lambda y: 0.5 + 0 * y, t_start=0, t_stop=5, nt=5)
    for method in ('BDF1', 'BDF2'):
        hb = O.Heat1D2PtsOracle(x_start=0, x_end=1, nx=1001, a=1, dtau=2 / 512, init_cond=C.heat_init, rhs=C.heat_rhs,
                                method=method, t_start=0, t_stop=2, nt=257)
        np.testing.assert_array_equal(hb.u0, g[f'heat1d2pts_{method}/in'])
        for k in range(3):
            dt = float(g[f'heat1d2pts_{method}/dt{k}'][0])
            got = hb.phi(hb.u0, 0.25, 0.25 + dt)
            ref = g[f'heat1d2pts_{method}/out{k}']
            assert np.max(np.abs(got - ref)) <= 1e-14 * np.max(np.abs(ref))
    np.testing.assert_array_equal(h2.u0, g['heat2d_65x49/in'])
    for k in range(3):
        dt = float(g[f'heat2d_65x49/dt{k}'][0])
        got = h2.phi(h2.u0, 0.1, 0.1 + dt)
        ref = g[f'heat2d_65x49/out{k}']
        assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))


def test_parallel_oracle_is_bitwise_serial():
    """The time-parallel CPU baseline (oracle/mgrit_oracle_mp.py: forked workers over shared level arrays, one block of
    intervals per worker like the reference's time ranks) reproduces the serial oracle exactly."""
    from oracle import mgrit_oracle_mp as OM
    from oracle_util import oracle_problem
    for name in ('heat1d_small_f_cf2', 'heat1d_varying_w', 'heat1d_small_jump'):
        case = C.CASES[name]
        serial = O.MgritOracle(oracle_problem(case), **case['solver'])
        ref = serial.solve()
        par = OM.ParallelMgritOracle(oracle_problem(case), workers=3, **case['solver'])
        try:
            got = par.solve()
            np.testing.assert_array_equal(got['conv'], ref['conv'])
            np.testing.assert_array_equal(par.u[0], serial.u[0])
            assert sum(par.nphi) == sum(serial.nphi)
        finally:
            par.close()
