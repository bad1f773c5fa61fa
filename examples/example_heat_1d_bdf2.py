"""examples/example_heat_1d_bdf2.py of the reference: pairs of time points, BDF2 on the fine grid over BDF1 on the
coarse grids (nt = 512 steps -> 257 pairs)."""
import numpy as np

from pymgrit_b200 import Heat1DBDF1, Heat1DBDF2, Mgrit

from _problems import rhs, init_cond


def build():
    t_stop, nt = 2, 512
    dtau = t_stop / nt
    kw = dict(x_start=0, x_end=1, nx=1001, a=1, dtau=dtau, rhs=rhs, init_cond=init_cond)
    heat0 = Heat1DBDF2(t_interval=np.linspace(0, t_stop, nt // 2 + 1), **kw)
    heat1 = Heat1DBDF1(t_interval=heat0.t[::2], **kw)
    heat2 = Heat1DBDF1(t_interval=heat1.t[::2], **kw)
    return dict(problem=[heat0, heat1, heat2])


if __name__ == '__main__':
    print(Mgrit(**build()).solve()['conv'])
