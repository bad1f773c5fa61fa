"""examples/example_dahlquist.py of the reference (README.rst:102-122): two-level MGRIT for u' = -u on [0, 5]."""
from pymgrit_b200 import Dahlquist, Mgrit, simple_setup_problem


def build():
    dahlquist = Dahlquist(t_start=0, t_stop=5, nt=101)
    return dict(problem=simple_setup_problem(problem=dahlquist, level=2, coarsening=2), tol=1e-10)


if __name__ == '__main__':
    info = Mgrit(**build()).solve()
    print(info['conv'])
