"""examples/example_spatial_coarsening.py of the reference: four levels, spatial coarsening by 2 on the first two level
transitions.  The reference's example defines its GridTransferHeat in Python; here the same transfer (full weighting /
linear interpolation) is the device class GridTransferHeat1D."""
from pymgrit_b200 import GridTransferCopy, GridTransferHeat1D, Heat1D, Mgrit

from _problems import rhs, init_cond


def build():
    kw = dict(x_start=0, x_end=2, a=1, rhs=rhs, init_cond=init_cond)
    heat0 = Heat1D(nx=2 ** 4 + 1, t_start=0, t_stop=2, nt=2 ** 7 + 1, **kw)
    heat1 = Heat1D(nx=2 ** 3 + 1, t_interval=heat0.t[::2], **kw)
    heat2 = Heat1D(nx=2 ** 2 + 1, t_interval=heat1.t[::2], **kw)
    heat3 = Heat1D(nx=2 ** 2 + 1, t_interval=heat2.t[::2], **kw)
    return dict(problem=[heat0, heat1, heat2, heat3],
                transfer=[GridTransferHeat1D(), GridTransferHeat1D(), GridTransferCopy()])


if __name__ == '__main__':
    print(Mgrit(**build()).solve()['conv'])
