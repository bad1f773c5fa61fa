"""Generate golden fixtures by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py [case ...]      # default: every case not marked slow
    python tests/golden/make_golden.py --all

Requires /root/reference (read-only) and the import stubs in oracle/stubs/ (mpi4py, matplotlib).
Writes tests/golden/<case>.npz with
    conv            residual history returned by Mgrit.solve()
    t_levels_n      number of time points per level
    u_rows_idx/u_rows   selected level-0 time points (all of them when small) and their values
    u_norms         np.linalg.norm of every level-0 time point
and tests/golden/decomposition.npz / phi_steps.npz (rank tables and single-Phi outputs).
Nothing here is imported by the product or run on the GPU box.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, 'oracle', 'stubs'), '/root/reference/src', os.path.join(ROOT, 'tests')]
warnings.filterwarnings('ignore')

from pymgrit.core.mgrit import Mgrit                      # noqa: E402
from pymgrit.core.at_mgrit import AtMgrit                 # noqa: E402
from pymgrit.heat.heat_1d import Heat1D                   # noqa: E402
from pymgrit.heat.heat_2d import Heat2D                   # noqa: E402
from pymgrit.heat.heat_1d_2pts_bdf1 import Heat1DBDF1     # noqa: E402
from pymgrit.heat.heat_1d_2pts_bdf2 import Heat1DBDF2     # noqa: E402
from pymgrit.advection.advection_1d import Advection1D    # noqa: E402
from pymgrit.dahlquist.dahlquist import Dahlquist         # noqa: E402
from pymgrit.brusselator.brusselator import Brusselator   # noqa: E402
from pymgrit.allen_cahn.allen_cahn import AllenCahn       # noqa: E402
import cases as C                                         # noqa: E402

APPS = {'heat1d': Heat1D, 'heat2d': Heat2D, 'advection1d': Advection1D, 'dahlquist': Dahlquist,
        'brusselator': Brusselator, 'allencahn': AllenCahn}


def heat1d2pts(method, **kw):
    return {'BDF1': Heat1DBDF1, 'BDF2': Heat1DBDF2}[method](**kw)


APPS['heat1d2pts'] = heat1d2pts


from pymgrit.core.grid_transfer import GridTransfer             # noqa: E402
from pymgrit.core.grid_transfer_copy import GridTransferCopy    # noqa: E402
from pymgrit.heat.heat_1d import VectorHeat1D                   # noqa: E402


class GridTransferHeat(GridTransfer):
    """User-side transfer class as a user of the reference writes it for examples/example_spatial_coarsening.py: full
    weighting down, linear interpolation up, on the interior points of Heat1D vectors."""

    def restriction(self, u):
        sol = u.get_values()
        out = np.zeros((len(sol) - 1) // 2)
        for i in range(len(out)):
            out[i] = sol[2 * i] * 1 / 4 + sol[2 * i + 1] * 1 / 2 + sol[2 * i + 2] * 1 / 4
        ret = VectorHeat1D(len(out))
        ret.set_values(out)
        return ret

    def interpolation(self, u):
        sol = u.get_values()
        out = np.zeros(len(sol) * 2 + 1)
        for i in range(len(sol)):
            out[i * 2] += 1 / 2 * sol[i]
            out[i * 2 + 1] += sol[i]
            out[i * 2 + 2] += 1 / 2 * sol[i]
        ret = VectorHeat1D(len(out))
        ret.set_values(out)
        return ret


def build_reference_transfer(case):
    if 'transfer' not in case:
        return None
    return [{'space': GridTransferHeat, 'copy': GridTransferCopy}[k]() for k in case['transfer']]


def build_reference_problem(case):
    grids = C.case_time_grids(case)
    return [APPS[case['app']](t_interval=t, **C.level_app_kw(case, l)) for l, t in enumerate(grids)]


def values_of(vec):
    vals = vec.get_values()
    if isinstance(vals, tuple):        # VectorHeat1D2Pts: (first, second, dtau)
        return np.stack([np.asarray(vals[0], dtype=float), np.asarray(vals[1], dtype=float)])
    return np.array(vals, dtype=float)


def run_case(name):
    case = C.CASES[name]
    problem = build_reference_problem(case)
    t0 = time.time()
    if 'at_k' in case:
        solver = AtMgrit(problem=problem, k=case['at_k'], logging_lvl=30, **case['solver'])
    else:
        solver = Mgrit(problem=problem, transfer=build_reference_transfer(case), logging_lvl=30, **case['solver'])
    info = solver.solve()
    wall = time.time() - t0
    u = [values_of(v) for v in solver.u[0]]
    n = len(u)
    per_point = int(np.size(u[0]))
    if n * per_point * 8 <= 256 * 1024:
        idx = np.arange(n)
    else:
        keep = max(4, min(n, (192 * 1024) // (per_point * 8)))
        idx = np.unique(np.round(np.linspace(0, n - 1, keep)).astype(int))
    out = dict(conv=np.asarray(info['conv']), t_levels_n=np.array([len(p.t) for p in problem]),
               u_rows_idx=idx, u_rows=np.stack([u[i] for i in idx]),
               u_norms=np.array([np.linalg.norm(x) for x in u]),
               ref_time_setup=info['time_setup'], ref_time_solve=info['time_solve'])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(f"{name}: {len(info['conv'])} iterations, conv[-1]={info['conv'][-1] if len(info['conv']) else None}, "
          f"{wall:.1f}s", flush=True)


def decomposition_tables():
    """Rank tables of setup_points_and_comm_info for faked ranks (how tests/core/test_mgrit.py:86-218 does it)."""
    out = {}
    configs = {'nt65_17_5_p7': ([65, 17, 5], 7), 'nt65_17_5_p4': ([65, 17, 5], 4), 'nt129_33_9_p3': ([129, 33, 9], 3),
               'nt33_17_9_5_p5': ([33, 17, 9, 5], 5), 'nt48_16_6_p4': ([48, 16, 6], 4), 'nt65_17_5_p1': ([65, 17, 5], 1),
               'nt257_65_17_p8': ([257, 65, 17], 8)}
    for key, (nts, size) in configs.items():
        if key == 'nt48_16_6_p4':
            t0 = np.linspace(0, 2, 48)
            ts = [t0, t0[::3], t0[::9]]
        else:
            ts = [np.linspace(0, 2, n) for n in nts]
        problem = [Dahlquist(t_interval=t) for t in ts]
        mg = Mgrit(problem=problem, nested_iteration=False, logging_lvl=30)
        for rank in range(size):
            mg.comm_time_size, mg.comm_time_rank = size, rank
            for attr in ('cpts', 'comm_front', 'comm_back', 'index_local_c', 'index_local_f', 'index_local',
                         'first_is_f_point', 'first_is_c_point', 'last_is_f_point', 'last_is_c_point', 'send_to',
                         'get_from', 'global_t'):
                setattr(mg, attr, [])
            mg.int_start = mg.int_stop = 0
            mg.t = [None] * len(ts)
            for lvl in range(len(ts)):
                mg.setup_points_and_comm_info(lvl=lvl)
            for lvl in range(len(ts)):
                pre = f'{key}/r{rank}/l{lvl}/'
                out[pre + 't'] = np.asarray(mg.t[lvl], dtype=float)
                out[pre + 'cpts'] = np.asarray(mg.cpts[lvl], dtype=int)
                out[pre + 'index_local'] = np.asarray(mg.index_local[lvl], dtype=int)
                out[pre + 'index_local_c'] = np.asarray(mg.index_local_c[lvl], dtype=int)
                out[pre + 'index_local_f'] = np.asarray(mg.index_local_f[lvl], dtype=int)
                out[pre + 'flags'] = np.array([mg.comm_front[lvl], mg.comm_back[lvl], mg.first_is_c_point[lvl],
                                               mg.first_is_f_point[lvl], mg.last_is_c_point[lvl],
                                               mg.last_is_f_point[lvl]], dtype=int)
                out[pre + 'send_get'] = np.array([mg.send_to[lvl], mg.get_from[lvl]], dtype=int)
    np.savez_compressed(os.path.join(HERE, 'decomposition.npz'), **out)
    print('decomposition tables:', len(out), 'arrays')


def phi_steps():
    """A few single applications of Phi at sizes the known-answer tests do not cover."""
    out = {}
    h = Heat1D(x_start=0, x_end=1, nx=1025, a=1, init_cond=C.heat_init, rhs=C.heat_rhs, t_start=0, t_stop=2, nt=16385)
    v = h.vector_t_start
    for k, dt in enumerate([2.0 / 16384, 8.0 / 16384, 32.0 / 16384, 2.0 / 2 ** 20, 0.5]):
        out[f'heat1d_1025/dt{k}'] = np.array([dt])
        out[f'heat1d_1025/out{k}'] = values_of(h.step(u_start=v, t_start=0.3, t_stop=0.3 + dt))
    a = Advection1D(c=1, x_start=-1, x_end=1, nx=4096, t_start=0, t_stop=2, nt=65537)
    v = a.vector_t_start
    for k, dt in enumerate([2.0 / 65536, 2.0 / 128, 0.25]):
        out[f'advection_4096/dt{k}'] = np.array([dt])
        out[f'advection_4096/out{k}'] = values_of(a.step(u_start=v, t_start=0.0, t_stop=dt))
    h2 = Heat2D(x_start=0, x_end=1, y_start=0, y_end=1, nx=65, ny=49, a=1, rhs=C.heat2d_rhs, init_cond=C.heat2d_init,
                bc_left=1.0, bc_top=lambda y: 0.5 + 0 * y, t_start=0, t_stop=5, nt=4097)
    v = h2.vector_t_start
    out['heat2d_65x49/in'] = values_of(v)
    for k, dt in enumerate([5.0 / 4096, 40.0 / 4096, 2.5]):
        out[f'heat2d_65x49/dt{k}'] = np.array([dt])
        out[f'heat2d_65x49/out{k}'] = values_of(h2.step(u_start=v, t_start=0.1, t_stop=0.1 + dt))
    for method, cls in (('BDF1', Heat1DBDF1), ('BDF2', Heat1DBDF2)):
        hb = cls(x_start=0, x_end=1, nx=1001, a=1, dtau=2 / 512, init_cond=C.heat_init, rhs=C.heat_rhs, t_start=0,
                 t_stop=2, nt=257)
        v = hb.vector_t_start
        out[f'heat1d2pts_{method}/in'] = values_of(v)
        for k, dt in enumerate([2.0 / 256, 8.0 / 256, 0.5]):
            out[f'heat1d2pts_{method}/dt{k}'] = np.array([dt])
            out[f'heat1d2pts_{method}/out{k}'] = values_of(hb.step(u_start=v, t_start=0.25, t_stop=0.25 + dt))
    np.savez_compressed(os.path.join(HERE, 'phi_steps.npz'), **out)
    print('phi steps:', len(out), 'arrays')


if __name__ == '__main__':
    args = sys.argv[1:]
    if args and args[0] != '--all':
        names = args
    else:
        names = [k for k, c in C.CASES.items() if args == ['--all'] or not c.get('slow')]
        decomposition_tables()
        phi_steps()
    for nm in names:
        run_case(nm)
