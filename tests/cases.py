"""Implementation-neutral registry of MGRIT test cases.

One case = application kind + its constructor arguments + how the time-grid hierarchy is built +
solver arguments.  The same case can be instantiated against
  * the unmodified reference (tests/golden/make_golden.py, build container only),
  * the CPU oracle (oracle/mgrit_oracle.py), and
  * the product package (pymgrit_b200, GPU),
so that parity tests read like the reference's own examples (examples/*.py, tests/mpi/*.py).
"""
import numpy as np


# -- problem data used by the reference's examples ------------------------------------------
def heat_rhs(x, t):          # examples/example_heat_1d.py:20-31
    return - np.sin(np.pi * x) * (np.sin(t) - 1 * np.pi ** 2 * np.cos(t))


def heat_init(x):            # examples/example_heat_1d.py:33-41
    return np.sin(np.pi * x)


def heat2d_rhs(x, y, t):     # docs/source/usage/parallelism.rst:121-125, examples/example_heat_2d.py
    return -np.sin(x) * np.sin(y) * (np.sin(t) - 2 * np.cos(t))


def heat2d_rhs_example(x, y, t):   # examples/example_heat_2d.py:36-47 (x_end = 0.75, y_end = 1.5, a = 3.5)
    return 5 * x * (0.75 - x) * y * (1.5 - y) + 10 * 3.5 * t * (y * (1.5 - y) + x * (0.75 - x))


def heat2d_rhs_xy(x, y, t):  # tests/heat/test_heat_2d.py:230-249
    return 2 * x * y


def heat2d_init(x, y):
    return np.sin(x) * np.sin(y)


def heat_rhs_nonsep(x, t):   # a right-hand side that is NOT a short sum of products X(x) T(t)
    return np.exp(-10 * (x - 0.5 - 0.2 * np.sin(t)) ** 2)


def heat_rhs_rank2(x, t):
    return np.sin(np.pi * x) * np.cos(t) + x * (1 - x) * t


HEAT = dict(x_start=0, x_end=1, a=1, init_cond=heat_init, rhs=heat_rhs)


def _simple(level, coarsening):
    return ('simple', level, coarsening)


# grids: ('simple', level, coarsening)  -> t = linspace(t_start, t_stop, nt), then t[::m] per level
#        ('nt', [nt0, nt1, ...])         -> each level its own linspace(t_start, t_stop, nt_l)
#        ('index', [idx1, idx2, ...])    -> level l+1 = level_l.t[idx_{l+1}] (slice or index array)
CASES = {
    # BASELINE.json configs[0] == README.rst:102-122 == examples/example_dahlquist.py
    'dahlquist_cfg1': dict(app='dahlquist', app_kw=dict(), t=(0, 5, 101), grids=_simple(2, 2),
                           solver=dict(tol=1e-10)),
    # tests/core/test_mgrit.py:59-70
    'heat1d_testmgrit': dict(app='heat1d', app_kw=dict(x_start=0, x_end=2, nx=5, a=1, rhs=heat_rhs,
                                                       init_cond=heat_init),
                             t=(0, 2, 65), grids=('nt', [65, 17, 5]),
                             solver=dict(cf_iter=1, nested_iteration=True, max_iter=2)),
    # tests/core/test_mgrit.py:72-84 (one level == time stepping)
    'heat1d_onelevel': dict(app='heat1d', app_kw=dict(x_start=0, x_end=2, nx=5, a=1, rhs=heat_rhs,
                                                      init_cond=heat_init),
                            t=(0, 2, 65), grids=('nt', [65]),
                            solver=dict(cf_iter=1, nested_iteration=True, max_iter=2)),
    # examples/example_heat_1d.py:43-53 -> tests/mpi/results/heat_1d
    'heat1d_example': dict(app='heat1d', app_kw=dict(nx=1001, **HEAT), t=(0, 2, 65),
                           grids=('nt', [65, 33, 17, 9, 5]),
                           solver=dict(cf_iter=1, cycle_type='F', nested_iteration=False, max_iter=10)),
    # examples/example_weighted_jacobi.py -> tests/mpi/results/weighted_jacobi (second solver)
    'heat1d_weighted': dict(app='heat1d', app_kw=dict(nx=1001, **HEAT), t=(0, 2, 65),
                            grids=('nt', [65, 33, 17, 9, 5]),
                            solver=dict(weight_c=1.3, tol=1e-8, cf_iter=1, cycle_type='F',
                                        nested_iteration=False, max_iter=10)),
    # BASELINE.json configs[1] at reduced nt (SURVEY.md section 8c sensitivity probe)
    'heat1d_cfg2_nt1025': dict(app='heat1d', app_kw=dict(nx=1025, **HEAT), t=(0, 2, 1025), grids=_simple(3, 4),
                               solver=dict(cf_iter=1, cycle_type='V', nested_iteration=True, tol=1e-10)),
    # BASELINE.json configs[1] in full
    'heat1d_cfg2': dict(app='heat1d', app_kw=dict(nx=1025, **HEAT), t=(0, 2, 16385), grids=_simple(3, 4),
                        solver=dict(cf_iter=1, cycle_type='V', nested_iteration=True, tol=1e-10), slow=True),
    # option coverage on a small heat problem
    'heat1d_small_v': dict(app='heat1d', app_kw=dict(nx=17, **HEAT), t=(0, 2, 129), grids=_simple(3, 4),
                           solver=dict(tol=1e-10)),
    'heat1d_small_f_cf2': dict(app='heat1d', app_kw=dict(nx=17, **HEAT), t=(0, 2, 129), grids=_simple(4, 2),
                               solver=dict(tol=1e-10, cycle_type='F', cf_iter=2)),
    'heat1d_small_cflist': dict(app='heat1d', app_kw=dict(nx=33, **HEAT), t=(0, 2, 65), grids=_simple(3, 2),
                                solver=dict(tol=1e-9, cf_iter=[2, 0, 1], nested_iteration=False)),
    'heat1d_small_tnorm1': dict(app='heat1d', app_kw=dict(nx=17, **HEAT), t=(0, 2, 65), grids=_simple(2, 4),
                                solver=dict(tol=1e-9, t_norm=1)),
    'heat1d_small_tnorminf': dict(app='heat1d', app_kw=dict(nx=17, **HEAT), t=(0, 2, 65), grids=_simple(2, 4),
                                  solver=dict(tol=1e-9, t_norm=3)),
    'heat1d_small_jump': dict(app='heat1d', app_kw=dict(nx=17, **HEAT), t=(0, 2, 65), grids=_simple(3, 2),
                              solver=dict(tol=1e-9, conv_crit=1)),
    'heat1d_small_weight': dict(app='heat1d', app_kw=dict(nx=65, **HEAT), t=(0, 2, 65), grids=_simple(3, 2),
                                solver=dict(tol=1e-9, weight_c=0.8, cycle_type='F')),
    # trailing F-points after the last C-point ((nt-1) % m != 0), and an odd spatial size
    'heat1d_trailing_f': dict(app='heat1d', app_kw=dict(nx=40, **HEAT), t=(0, 1, 48), grids=_simple(3, 3),
                              solver=dict(tol=1e-9)),
    # non-uniform time grid (dt varies per step)
    'heat1d_nonuniform_t': dict(app='heat1d', app_kw=dict(nx=33, **HEAT),
                                t_interval=np.linspace(0, 1, 65) ** 1.5 * 2, grids=_simple(3, 2),
                                solver=dict(tol=1e-9)),
    # non-uniform coarsening with adjacent C-points (tests/mpi/varying_coarsening.py, on the PDE kernels)
    'heat1d_varying': dict(app='heat1d', app_kw=dict(nx=40, **HEAT), t=(0, 2, 65),
                           grids=('index', [np.array([0, 3, 4, 8, 12, 13, 14, 20, 25, 26, 32, 36, 40, 41, 47, 50, 55, 56,
                                                      60, 64]), slice(None, None, 2)]),
                           solver=dict(tol=1e-7, nested_iteration=False)),
    'heat1d_varying_w': dict(app='heat1d', app_kw=dict(nx=40, **HEAT), t=(0, 2, 65),
                             grids=('index', [np.array([0, 1, 2, 8, 12, 13, 14, 20, 25, 26, 32, 36, 40, 41, 47, 50, 55,
                                                        56, 60, 64]), slice(None, None, 3)]),
                             solver=dict(tol=1e-7, weight_c=1.2, cycle_type='F')),
    # right-hand sides: rank-2 separable, and not separable (dense table path)
    'heat1d_rhs_rank2': dict(app='heat1d', app_kw=dict(x_start=0, x_end=1, nx=65, a=0.5, init_cond=heat_init,
                                                       rhs=heat_rhs_rank2),
                             t=(0, 2, 65), grids=_simple(2, 4), solver=dict(tol=1e-9)),
    'heat1d_rhs_nonsep': dict(app='heat1d', app_kw=dict(x_start=0, x_end=1, nx=65, a=0.1, init_cond=heat_init,
                                                        rhs=heat_rhs_nonsep),
                              t=(0, 2, 65), grids=_simple(2, 4), solver=dict(tol=1e-9)),
    'heat1d_zero_rhs': dict(app='heat1d', app_kw=dict(x_start=0, x_end=1, nx=129, a=1, init_cond=heat_init),
                            t=(0, 0.5, 33), grids=_simple(2, 2), solver=dict(tol=1e-9)),
    # larger spatial size -> multi-warp team per system
    'heat1d_nx4097': dict(app='heat1d', app_kw=dict(nx=4097, **HEAT), t=(0, 2, 129), grids=_simple(2, 4),
                          solver=dict(tol=1e-8)),
    # examples/example_multilevel_structure.py -> tests/mpi/results/multilevel_structure
    'dahlquist_ml1': dict(app='dahlquist', app_kw=dict(), t=(0, 5, 101), grids=_simple(3, 2), solver=dict(tol=1e-10)),
    'dahlquist_ml2': dict(app='dahlquist', app_kw=dict(), t=(0, 5, 101), grids=('nt', [101, 51, 26]),
                          solver=dict(tol=1e-10)),
    'dahlquist_ml3': dict(app='dahlquist', app_kw=dict(), t_interval=np.linspace(0, 5, 101),
                          grids=('index', [slice(None, None, 2), slice(None, None, 2)]), solver=dict(tol=1e-10)),
    # tests/mpi/varying_coarsening.py -> tests/mpi/results/varying_coarsening
    'dahlquist_varying': dict(app='dahlquist', app_kw=dict(), t=(0, 5, 65),
                              grids=('index', [np.array([0, 3, 10, 12, 14, 17, 23, 27, 33, 34, 55, 57, 59, 61, 63, 64]),
                                               slice(None, None, 2), slice(None, None, 2), slice(None, None, 2)]),
                              solver=dict(tol=1e-10, nested_iteration=False)),
    # tests/mpi/procs_without_points.py -> tests/mpi/results/procs_without_points
    'dahlquist_procs_wo_pts': dict(app='dahlquist', app_kw=dict(), t=(0, 5, 129),
                                   grids=('index', [slice(None, None, 16), slice(None, None, 2),
                                                    slice(None, None, 2), slice(None, None, 2)]),
                                   solver=dict(tol=1e-10)),
    # examples/example_time_integrators.py -> tests/mpi/results/time_integrators (method differs per level)
    'dahlquist_integrators': dict(app='dahlquist', app_kw=[dict(method='MR'), dict(method='BE')], t=(0, 5, 101),
                                  grids=('nt', [101, 51]), solver=dict()),
    'dahlquist_tr_fe': dict(app='dahlquist', app_kw=[dict(method='TR'), dict(method='FE', constant_lambda=-0.5)],
                            t=(0, 5, 101), grids=('nt', [101, 51]), solver=dict(tol=1e-9)),
    # examples/example_brusselator.py -> tests/mpi/results/brusselator
    'brusselator_example': dict(app='brusselator', app_kw=dict(), t=(0, 12, 641),
                                grids=('index', [slice(None, None, 20)]), solver=dict(cf_iter=1)),
    # examples/example_allen_cahn.py:36-37 (IMEX, two levels) at nx = 32; F-cycle / three levels / jump criterion variants
    'allencahn_example': dict(app='allencahn', app_kw=dict(nx=32, method='IMEX'), t=(0, 0.032, 33),
                              grids=_simple(2, 2), solver=dict(tol=1e-9)),
    'allencahn_3lvl_f': dict(app='allencahn', app_kw=dict(nx=24, method='IMEX', nu=2, eps=0.05), t=(0, 0.032, 65),
                             grids=_simple(3, 4), solver=dict(tol=1e-9, cycle_type='F', cf_iter=2, nested_iteration=False)),
    'allencahn_jump': dict(app='allencahn', app_kw=dict(nx=15, method='IMEX', nu=2, eps=0.06), t=(0, 0.02, 41),
                           grids=_simple(2, 4), solver=dict(tol=1e-8, conv_crit=1)),
    # examples/example_advection.py
    'advection_example': dict(app='advection1d', app_kw=dict(c=1, x_start=-1, x_end=1, nx=129), t=(0, 2, 129),
                              grids=('nt', [129, 65]), solver=dict(cf_iter=1, nested_iteration=False)),
    # BASELINE.json configs[3] at reduced size
    'advection_cfg4_small': dict(app='advection1d', app_kw=dict(c=1, x_start=-1, x_end=1, nx=257), t=(0, 2, 1025),
                                 grids=_simple(5, 2), solver=dict(tol=1e-10, cf_iter=1, nested_iteration=True)),
    # two levels with a long coarsest level (513 points): the coarsest solve runs in Fourier space (csrc/fourier.cu); nx - 1
    # = 256 unknowns, and 301 - 1 = 300 = 2^2 3 5^2 (a length that is not a power of two: Bluestein's transform)
    'advection_two_level_fourier': dict(app='advection1d', app_kw=dict(c=1, x_start=-1, x_end=1, nx=257), t=(0, 2, 1025),
                                        grids=_simple(2, 2), solver=dict(tol=1e-10, cf_iter=1, nested_iteration=True)),
    'advection_nx301_fourier': dict(app='advection1d', app_kw=dict(c=1, x_start=-1, x_end=1, nx=301), t=(0, 1, 513),
                                    grids=_simple(3, 2), solver=dict(tol=1e-9, cf_iter=1, cycle_type='F')),
    'advection_nx4096_short': dict(app='advection1d', app_kw=dict(c=1, x_start=-1, x_end=1, nx=4096), t=(0, 2 / 256, 257),
                                   grids=_simple(3, 4), solver=dict(tol=1e-10)),
    # examples/example_heat_2d.py -> tests/mpi/results/heat_2d
    'heat2d_example': dict(app='heat2d', app_kw=dict(x_start=0, x_end=0.75, y_start=0, y_end=1.5,
                                                     nx=55, ny=125, a=3.5, rhs=heat2d_rhs_example),
                           t=(0, 1, 33), grids=('index', [slice(None, None, 2)]), solver=dict(cycle_type='V')),
    # BASELINE.json configs[2] at reduced size
    'heat2d_cfg3_small': dict(app='heat2d', app_kw=dict(x_start=0, x_end=1, y_start=0, y_end=1, nx=33, ny=33, a=1,
                                                        rhs=heat2d_rhs),
                              t=(0, 5, 513), grids=_simple(3, 8), solver=dict(tol=1e-10, cycle_type='F')),
    'heat2d_bc': dict(app='heat2d', app_kw=dict(x_start=0, x_end=1, y_start=3, y_end=4, nx=17, ny=21, a=1,
                                                rhs=heat2d_rhs_xy, init_cond=heat2d_init, bc_left=2.0, bc_right=1.0,
                                                bc_bottom=0.5, bc_top=1.5),
                      t=(0, 1, 65), grids=_simple(2, 4), solver=dict(tol=1e-8)),
    # two-point (pair) states, heat/heat_1d_2pts_bdf{1,2}.py: examples/example_heat_1d_bdf2.py (BDF2 on the fine grid,
    # BDF1 on the coarse grids; nt = 512 steps -> 257 pairs, dtau = t_stop / nt), and smaller variants
    'heat1d_bdf2_example': dict(app='heat1d2pts',
                                app_kw=[dict(nx=1001, dtau=2 / 512, method='BDF2', **HEAT),
                                        dict(nx=1001, dtau=2 / 512, method='BDF1', **HEAT),
                                        dict(nx=1001, dtau=2 / 512, method='BDF1', **HEAT)],
                                t=(0, 2, 257), grids=_simple(3, 2), solver=dict()),
    'heat1d_bdf1_small': dict(app='heat1d2pts', app_kw=dict(nx=17, dtau=2 / 128, method='BDF1', **HEAT),
                              t=(0, 2, 65), grids=_simple(3, 2), solver=dict(tol=1e-9)),
    'heat1d_bdf2_small_f': dict(app='heat1d2pts', app_kw=dict(nx=34, dtau=1 / 256, method='BDF2', **HEAT),
                                t=(0, 1, 129), grids=_simple(3, 4), solver=dict(tol=1e-9, cycle_type='F')),
    'heat1d_bdf2_nonuniform': dict(app='heat1d2pts',
                                   app_kw=[dict(nx=65, dtau=0.004, method='BDF2', x_start=0, x_end=1, a=0.5,
                                                init_cond=heat_init, rhs=heat_rhs_rank2),
                                           dict(nx=65, dtau=0.004, method='BDF1', x_start=0, x_end=1, a=0.5,
                                                init_cond=heat_init, rhs=heat_rhs_rank2)],
                                   t_interval=0.01 + np.linspace(0, 1, 49) ** 1.3 * 1.5, grids=_simple(2, 3),
                                   solver=dict(tol=1e-9, weight_c=0.9, nested_iteration=False)),
    # examples/example_spatial_coarsening.py -> tests/mpi/results/spatial_coarsening: spatial coarsening by 2 on the
    # first two level transitions (full weighting / linear interpolation), identity on the last
    'heat1d_spatial_example': dict(app='heat1d',
                                   app_kw=[dict(x_start=0, x_end=2, nx=17, a=1, rhs=heat_rhs, init_cond=heat_init),
                                           dict(x_start=0, x_end=2, nx=9, a=1, rhs=heat_rhs, init_cond=heat_init),
                                           dict(x_start=0, x_end=2, nx=5, a=1, rhs=heat_rhs, init_cond=heat_init),
                                           dict(x_start=0, x_end=2, nx=5, a=1, rhs=heat_rhs, init_cond=heat_init)],
                                   t=(0, 2, 129), grids=_simple(4, 2), transfer=['space', 'space', 'copy'],
                                   solver=dict()),
    # larger grids (several lanes / warps per system), F-cycles, FCF twice, trailing F-points, tighter tolerance
    'heat1d_spatial_large': dict(app='heat1d',
                                 app_kw=[dict(nx=1025, **HEAT), dict(nx=513, **HEAT), dict(nx=513, **HEAT)],
                                 t=(0, 2, 98), grids=_simple(3, 3), transfer=['space', 'copy'],
                                 solver=dict(tol=1e-6, cycle_type='F', cf_iter=2)),
    'heat1d_spatial_nonested': dict(app='heat1d',
                                    app_kw=[dict(nx=65, **HEAT), dict(nx=65, **HEAT), dict(nx=33, **HEAT)],
                                    t=(0, 1, 65), grids=_simple(3, 4), transfer=['copy', 'space'],
                                    solver=dict(tol=1e-7, nested_iteration=False, weight_c=1.1)),
    # AT-MGRIT (core/at_mgrit.py) on one time rank: tests/core/test_at_mgrit.py:33-45 (k = 2), and longer local grids
    'heat1d_atmgrit_test': dict(app='heat1d', app_kw=dict(x_start=0, x_end=2, nx=5, a=1, rhs=heat_rhs,
                                                          init_cond=heat_init),
                                t=(0, 2, 65), grids=('nt', [65, 17, 5]), at_k=2,
                                solver=dict(cf_iter=1, nested_iteration=False, max_iter=2)),
    'heat1d_atmgrit_k8': dict(app='heat1d', app_kw=dict(nx=129, **HEAT), t=(0, 2, 513), grids=_simple(3, 4), at_k=8,
                              solver=dict(tol=1e-8, nested_iteration=True)),
    'heat1d_atmgrit_k5_f': dict(app='heat1d', app_kw=dict(nx=33, **HEAT), t=(0, 1, 257), grids=_simple(3, 4), at_k=5,
                                solver=dict(tol=1e-8, cycle_type='F', conv_crit=1, nested_iteration=False)),
    # local convergence criteria (mgrit.py:434-454) on one time rank
    'heat1d_small_local_res': dict(app='heat1d', app_kw=dict(nx=17, **HEAT), t=(0, 2, 65), grids=_simple(3, 2),
                                   solver=dict(tol=1e-9, conv_crit=2)),
    'heat1d_small_local_jump': dict(app='heat1d', app_kw=dict(nx=17, **HEAT), t=(0, 2, 65), grids=_simple(3, 2),
                                    solver=dict(tol=1e-9, conv_crit=3, t_norm=3)),
    # heat_2d.py:341-366: Crank-Nicolson (theta = 1/2) with non-zero Dirichlet data, and forward Euler (theta = 0)
    'heat2d_cn': dict(app='heat2d', app_kw=dict(x_start=0, x_end=1, y_start=0, y_end=1, nx=17, ny=13, a=1,
                                                rhs=heat2d_rhs, init_cond=heat2d_init, method='CN', bc_left=1.0,
                                                bc_top=0.5),
                      t=(0, 1, 33), grids=_simple(2, 4), solver=dict(tol=1e-9)),
    'heat2d_cn_3lvl': dict(app='heat2d', app_kw=dict(x_start=0, x_end=0.75, y_start=0, y_end=1.5, nx=15, ny=25, a=0.5,
                                                     rhs=heat2d_rhs, init_cond=heat2d_init, method='CN',
                                                     bc_bottom=lambda y: 0.25 * y),
                           t=(0, 1, 65), grids=_simple(3, 4), solver=dict(tol=1e-9, cycle_type='F')),
    'heat2d_fe': dict(app='heat2d', app_kw=dict(x_start=0, x_end=1, y_start=0, y_end=1, nx=9, ny=9, a=1,
                                                rhs=heat2d_rhs, init_cond=heat2d_init, method='FE'),
                      t=(0, 0.05, 33), grids=_simple(2, 2), solver=dict(tol=1e-9)),
}


def case_time_grids(case):
    """List of time grids, one per level, exactly as the reference examples build them."""
    if 't_interval' in case:
        t0 = np.asarray(case['t_interval'], dtype=float)
        rng = (t0[0], t0[-1])
    else:
        a, b, nt = case['t']
        t0 = np.linspace(a, b, nt)
        rng = (a, b)
    kind = case['grids'][0]
    grids = [t0]
    if kind == 'simple':
        _, level, m = case['grids']
        for _ in range(level - 1):
            grids.append(grids[-1][::m])
    elif kind == 'nt':
        grids = [np.linspace(rng[0], rng[1], n) for n in case['grids'][1]]
    elif kind == 'index':
        for idx in case['grids'][1]:
            grids.append(grids[-1][idx])
    else:
        raise ValueError(kind)
    return grids


def level_app_kw(case, lvl):
    kw = case['app_kw']
    return dict(kw[lvl] if isinstance(kw, list) else kw)
