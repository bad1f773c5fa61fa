"""examples/example_allen_cahn.py of the reference: two-level MGRIT for the 2-D Allen-Cahn equation, IMEX time stepping
(the application runs on the batched path: pymgrit_b200/core/batched.py)."""
from pymgrit_b200 import AllenCahn, Mgrit


def build():
    problem_level_0 = AllenCahn(t_start=0, t_stop=0.032, nt=33, method='IMEX')
    problem_level_1 = AllenCahn(t_interval=problem_level_0.t[::2], method='IMEX')
    return dict(problem=[problem_level_0, problem_level_1], tol=1e-9)


if __name__ == '__main__':
    solver = Mgrit(**build())
    info = solver.solve()
    app = solver.problem[0]
    last = solver.u[0][-1]
    print(info['conv'], 'radius', app.compute_radius(last), 'exact', app.exact_radius(solver.t[0][-1]))
