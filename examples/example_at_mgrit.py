"""AT-MGRIT as in tests/core/test_at_mgrit.py:33-45 and examples/at_mgrit/runme_heat1d_m_k.py of the reference: local
coarse grids of k points on the coarsest level instead of the sequential coarse solve."""
from pymgrit_b200 import AtMgrit, Heat1D

from _problems import rhs, init_cond


def build():
    kw = dict(x_start=0, x_end=2, nx=5, a=1, rhs=rhs, init_cond=init_cond, t_start=0, t_stop=2)
    return dict(problem=[Heat1D(nt=nt, **kw) for nt in (65, 17, 5)], k=2, cf_iter=1, nested_iteration=False, max_iter=2)


if __name__ == '__main__':
    print(AtMgrit(**build()).solve()['conv'])      # the reference's test expects [0.1767778, 0.01223507]
