"""examples/example_heat_2d.py of the reference: two-level V-cycles for the 2-D heat equation on [0, 0.75] x [0, 1.5],
with an output function that collects the solution at the last time point."""
import numpy as np

from pymgrit_b200 import Heat2D, Mgrit


def rhs(x, y, t):
    return 5 * x * (0.75 - x) * y * (1.5 - y) + 10 * 3.5 * t * (y * (1.5 - y) + x * (0.75 - x))


def build(collect=None):
    heat0 = Heat2D(x_start=0, x_end=0.75, y_start=0, y_end=1.5, nx=55, ny=125, a=3.5, rhs=rhs, t_start=0, t_stop=1, nt=33)
    heat1 = Heat2D(x_start=0, x_end=0.75, y_start=0, y_end=1.5, nx=55, ny=125, a=3.5, rhs=rhs, t_interval=heat0.t[::2])

    def output_fcn(self):
        if collect is not None:
            collect.append(np.array(self.u[0][-1].get_values()))

    return dict(problem=[heat0, heat1], cycle_type='V', output_fcn=output_fcn)


if __name__ == '__main__':
    got = []
    info = Mgrit(**build(got)).solve()
    print(info['conv'], got[-1].shape)
