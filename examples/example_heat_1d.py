"""examples/example_heat_1d.py of the reference: five-level F-cycles with FCF-relaxation for the 1-D heat equation,
each level with its own time grid."""
from pymgrit_b200 import Heat1D, Mgrit

from _problems import rhs, init_cond


def build():
    kw = dict(x_start=0, x_end=1, nx=1001, a=1, init_cond=init_cond, rhs=rhs, t_start=0, t_stop=2)
    problem = [Heat1D(nt=nt, **kw) for nt in (65, 33, 17, 9, 5)]
    return dict(problem=problem, max_iter=10, tol=1e-7, nested_iteration=False, cf_iter=1, cycle_type='F')


if __name__ == '__main__':
    solver = Mgrit(**build())
    info = solver.solve()
    print(info['conv'], solver.u[0][-1].get_values()[:3])
