"""GPU tests of the plugin API itself (run with -m gpu): Application.step and the Vector operations of every device
application, against the reference's own known-answer vectors (tests/heat/test_heat_1d.py, test_heat_2d.py,
tests/advection/test_advection_1d.py, tests/dahlquist/test_dahlquist.py, tests/brusselator/test_brusselator.py) and the
single-Phi fixtures produced by the unmodified reference (tests/golden/phi_steps.npz).  Every call goes through
pymgrit_b200 -> C ABI -> kernels."""
import copy

import numpy as np
import pytest

import cases as C
from oracle_util import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def P():
    import pymgrit_b200
    return pymgrit_b200


def test_heat1d_step_known_answer(P):            # tests/heat/test_heat_1d.py:31-42
    p = P.Heat1D(a=1, init_cond=lambda x: 2 * x, x_start=0, x_end=1, nx=6, t_start=0, t_stop=1, nt=11)
    out = p.step(u_start=p.vector_t_start, t_start=0, t_stop=0.1)
    np.testing.assert_almost_equal(out.get_values(), np.array([0.28164, 0.51593599, 0.63660638, 0.53191933]))


def test_heat2d_step_known_answer(P):            # tests/heat/test_heat_2d.py:230-249
    p = P.Heat2D(a=1, x_start=0, x_end=1, y_start=3, y_end=4, nx=5, ny=5, rhs=lambda x, y, t: 2 * x * y,
                 t_start=0, t_stop=1, nt=11)
    out = p.step(u_start=p.vector_t_start, t_start=0, t_stop=0.1)
    want = np.array([[0., 0., 0., 0., 0.], [0., 0.06659024, 0.08719337, 0.07227713, 0.],
                     [0., 0.11922399, 0.15502696, 0.12990086, 0.], [0., 0.12666875, 0.16193148, 0.1391124, 0.],
                     [0., 0., 0., 0., 0.]])
    np.testing.assert_almost_equal(out.get_values(), want)


def test_heat2d_constructor_errors(P):           # tests/heat/test_heat_2d.py:200-227 (argument checks)
    kw = dict(a=1, x_start=0, x_end=1, y_start=3, y_end=4, nx=5, ny=5, t_start=0, t_stop=1, nt=11)
    with pytest.raises(Exception):
        P.Heat2D(method='unknown', **kw)
    with pytest.raises(Exception):
        P.Heat2D(bc_left='0', **kw)
    with pytest.raises(Exception):
        P.Heat2D(method='FE', bc_top=1.0, **kw)  # FE with non-zero Dirichlet data is not reproduced: fails loudly
    assert P.Heat2D(method='CN', **kw).theta == 0.5


def test_advection_step_known_answer(P):         # tests/advection/test_advection_1d.py:33-45
    p = P.Advection1D(c=1, x_start=0, x_end=1, nx=6, t_start=0, t_stop=1, nt=11)
    out = p.step(u_start=p.vector_t_start, t_start=0, t_stop=0.1)
    np.testing.assert_almost_equal(out.get_values(), np.array([0.868043, 0.92987396, 0.87805385, 0.75780217, 0.604129]))


def test_dahlquist_step_known_answer(P):         # tests/dahlquist/test_dahlquist.py:55-62
    p = P.Dahlquist(t_start=0, t_stop=1, nt=11)
    out = p.step(u_start=p.vector_t_start, t_start=0, t_stop=0.1)
    np.testing.assert_almost_equal(out.get_values(), 0.9090909090909091)


def test_brusselator_step_known_answer(P):       # tests/brusselator/test_brusselator.py:24-31
    p = P.Brusselator(t_start=0, t_stop=1, nt=11)
    out = p.step(u_start=p.vector_template.clone_zero(), t_start=0, t_stop=0.1)
    np.testing.assert_almost_equal(out.get_values(), np.array([0.08240173, 0.01319825]))


def test_heat1d_two_point_step_known_answers(P):
    # tests/heat/test_heat_1d_2pts_bdf1.py:35-54 (dt = dtau: the first of the two solves is the identity)
    p = P.Heat1DBDF1(a=1, init_cond=lambda x: 2 * x, x_start=0, x_end=1, nx=11, dtau=0.1, t_start=0, t_stop=1, nt=11)
    out = p.step(u_start=p.vector_t_start, t_start=0, t_stop=0.1)
    assert isinstance(out, P.VectorHeat1D2Pts)
    first, second, dtau = out.get_values()
    np.testing.assert_almost_equal(first, np.array(
        [0.14498001, 0.28445802, 0.41238183, 0.52154382, 0.6028602, 0.6444626, 0.63051125, 0.53961104, 0.34267192]))
    np.testing.assert_almost_equal(second, np.array(
        [0.08691756, 0.16802887, 0.23749726, 0.2894772, 0.31825048, 0.31856279, 0.28628511, 0.21958482, 0.12088191]))
    assert dtau == 0.1
    # tests/heat/test_heat_1d_2pts_bdf2.py:12-62
    z = P.Heat1DBDF2(a=1, x_start=0, x_end=1, nx=11, dtau=0.1, t_start=0, t_stop=1, nt=11)
    assert z.nx == 9 and isinstance(z.vector_template, P.VectorHeat1D2Pts)
    np.testing.assert_equal(z.vector_t_start.get_values()[0], np.zeros(9))
    np.testing.assert_equal(z.vector_t_start.get_values()[1], np.zeros(9))
    p = P.Heat1DBDF2(a=1, init_cond=lambda x: 2 * x, x_start=0, x_end=1, nx=11, dtau=0.1, t_start=0, t_stop=1, nt=5)
    np.testing.assert_almost_equal(p.vector_t_start.get_values()[0], np.array([0.2, 0.4, 0.6, 0.8, 1., 1.2, 1.4, 1.6, 1.8]))
    np.testing.assert_almost_equal(p.vector_t_start.get_values()[1], np.array(
        [0.15656217, 0.30443677, 0.43319873, 0.52860043, 0.56972221, 0.52478844, 0.34481236, -0.04620125, -0.76645512]))
    first, second, dtau = p.step(u_start=p.vector_t_start, t_start=0, t_stop=0.2).get_values()
    np.testing.assert_almost_equal(first, np.array(
        [0.07115547, 0.13167183, 0.17105162, 0.1794494, 0.1490445, 0.07705183, -0.02834074, -0.1369469, -0.17685485]))
    np.testing.assert_almost_equal(second, np.array(
        [0.01235156, 0.02015287, 0.01986458, 0.01000559, -0.00781242, -0.02812508, -0.04182745, -0.03889518,
         -0.01671786]))


def test_heat1d_two_point_vector(P):
    """tests/heat/test_vector_heat_1d_2pts.py: the pair vector's operations."""
    v = P.VectorHeat1D2Pts(3, 0.1)
    v.set_values(np.array([1.0, 2, 3]), np.array([4.0, 5, 6]), 0.1)
    w = P.VectorHeat1D2Pts(3, 0.1)
    w.set_values(np.ones(3), 2 * np.ones(3), 0.1)
    s, d, m = v + w, v - w, v * 3
    np.testing.assert_array_equal(s.get_values()[0], [2, 3, 4])
    np.testing.assert_array_equal(s.get_values()[1], [6, 7, 8])
    np.testing.assert_array_equal(d.get_values()[1], [2, 3, 4])
    np.testing.assert_array_equal(m.get_values()[0], [3, 6, 9])
    assert s.get_values()[2] == 0.1 and v.size == 3
    assert abs(v.norm() - np.linalg.norm([1, 2, 3, 4, 5, 6])) < 1e-14
    c = v.clone()
    c.unpack(np.array([[9.0, 9, 9], [8.0, 8, 8]]))
    np.testing.assert_array_equal(v.pack(), np.array([[1.0, 2, 3], [4.0, 5, 6]]))
    np.testing.assert_array_equal(c.get_values()[1], [8, 8, 8])
    assert not np.any(v.clone_zero().get_values()[0]) and v.clone_rand().get_values()[1].shape == (3,)
    v2 = copy.deepcopy(v)
    np.testing.assert_array_equal(v2.get_values()[0], [1, 2, 3])


def test_grid_transfer_heat1d_vectors(P):
    """GridTransferHeat1D vector by vector against the arithmetic of examples/example_spatial_coarsening.py:33-79."""
    tr = P.GridTransferHeat1D()
    rng = np.random.default_rng(3)
    sol = rng.standard_normal(15)
    u = P.VectorHeat1D(15)
    u.set_values(sol)
    want = np.array([sol[2 * i] * 1 / 4 + sol[2 * i + 1] * 1 / 2 + sol[2 * i + 2] * 1 / 4 for i in range(7)])
    r = tr.restriction(u)
    assert isinstance(r, P.VectorHeat1D)
    np.testing.assert_array_equal(r.get_values(), want)
    back = np.zeros(15)
    for i in range(7):
        back[2 * i] += 1 / 2 * want[i]
        back[2 * i + 1] += want[i]
        back[2 * i + 2] += 1 / 2 * want[i]
    np.testing.assert_array_equal(tr.interpolation(r).get_values(), back)
    c = P.GridTransferCopy()
    np.testing.assert_array_equal(c.restriction(u).get_values(), sol)
    assert P.GridTransferHeat is P.GridTransferHeat1D


def test_phi_step_fixtures(P):
    """Single Phi at BASELINE sizes against the unmodified reference (1e-10 relative)."""
    g = load_golden('phi_steps')
    h = P.Heat1D(x_start=0, x_end=1, nx=1025, a=1, init_cond=C.heat_init, rhs=C.heat_rhs, t_start=0, t_stop=2, nt=5)
    for k in range(5):
        dt = float(g[f'heat1d_1025/dt{k}'][0])
        got = h.step(u_start=h.vector_t_start, t_start=0.3, t_stop=0.3 + dt).get_values()
        ref = g[f'heat1d_1025/out{k}']
        assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref))
    a = P.Advection1D(c=1, x_start=-1, x_end=1, nx=4096, t_start=0, t_stop=2, nt=5)
    for k in range(3):
        dt = float(g[f'advection_4096/dt{k}'][0])
        got = a.step(u_start=a.vector_t_start, t_start=0.0, t_stop=dt).get_values()
        ref = g[f'advection_4096/out{k}']
        assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref))
    for method, cls in (('BDF1', P.Heat1DBDF1), ('BDF2', P.Heat1DBDF2)):
        hb = cls(x_start=0, x_end=1, nx=1001, a=1, dtau=2 / 512, init_cond=C.heat_init, rhs=C.heat_rhs, t_start=0,
                 t_stop=2, nt=257)
        start = g[f'heat1d2pts_{method}/in']
        first, second, _ = hb.vector_t_start.get_values()
        assert np.max(np.abs(np.stack([first, second]) - start)) <= 1e-10 * np.max(np.abs(start))
        for k in range(3):
            dt = float(g[f'heat1d2pts_{method}/dt{k}'][0])
            first, second, _ = hb.step(u_start=hb.vector_t_start, t_start=0.25, t_stop=0.25 + dt).get_values()
            ref = g[f'heat1d2pts_{method}/out{k}']
            assert np.max(np.abs(np.stack([first, second]) - ref)) <= 1e-10 * np.max(np.abs(ref))
    h2 = P.Heat2D(x_start=0, x_end=1, y_start=0, y_end=1, nx=65, ny=49, a=1, rhs=C.heat2d_rhs, init_cond=C.heat2d_init,
                  bc_left=1.0, bc_top=lambda y: 0.5 + 0 * y, t_start=0, t_stop=5, nt=5)
    np.testing.assert_array_equal(h2.vector_t_start.get_values(), g['heat2d_65x49/in'])
    for k in range(3):
        dt = float(g[f'heat2d_65x49/dt{k}'][0])
        got = h2.step(u_start=h2.vector_t_start, t_start=0.1, t_stop=0.1 + dt).get_values()
        ref = g[f'heat2d_65x49/out{k}']
        assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_heat2d_512_step_against_oracle(P):
    """BASELINE.json configs[2] size (512 x 512): one backward-Euler step on level 0 and one with the coarsest
    level's dt, GPU (sine space) against the CPU oracle (sparse direct solve of the same system)."""
    from oracle import mgrit_oracle as O
    kw = dict(x_start=0, x_end=1, y_start=0, y_end=1, nx=512, ny=512, a=1, rhs=C.heat2d_rhs, init_cond=C.heat2d_init)
    app = P.Heat2D(t_start=0, t_stop=5, nt=4097, **kw)
    orc = O.Heat2DOracle(t_start=0, t_stop=5, nt=4097, **kw)
    for dt in (5.0 / 4096, 512 * 5.0 / 4096):
        got = app.step(u_start=app.vector_t_start, t_start=0.5, t_stop=0.5 + dt).get_values()
        ref = orc.phi(orc.u0, 0.5, 0.5 + dt)
        assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref))
    # the transforms are inverse to each other and orthogonal
    fam = app.family()
    v = app.vector_t_start
    rows = fam.to_rows(v.device_values.reshape(1, 512, 512))
    back = fam.from_rows(rows)[0].cpu().numpy()
    assert np.max(np.abs(back - v.get_values())) <= 1e-13
    assert abs(float(rows.norm()) - v.norm()) <= 1e-12 * v.norm()


@pytest.mark.parametrize('make', [
    lambda P: P.Heat1D(a=1, init_cond=lambda x: 2 * x, x_start=0, x_end=1, nx=12, t_start=0, t_stop=1, nt=11),
    lambda P: P.Heat2D(a=1, x_start=0, x_end=1, y_start=3, y_end=4, nx=7, ny=9, init_cond=lambda x, y: x * y,
                       t_start=0, t_stop=1, nt=11),
    lambda P: P.Advection1D(c=1, x_start=0, x_end=1, nx=9, t_start=0, t_stop=1, nt=11),
    lambda P: P.Brusselator(t_start=0, t_stop=1, nt=11),
], ids=['heat1d', 'heat2d', 'advection1d', 'brusselator'])
def test_vector_operations(P, make):
    """The Vector contract (core/vector.py:38-151): out-of-place +, -, *, norm, clone family, set/get, pack/unpack."""
    app = make(P)
    v = app.vector_t_start
    a = np.array(v.get_values(), dtype=float)
    w = v.clone_rand()
    b = np.array(w.get_values(), dtype=float)
    assert a.shape == b.shape == tuple(app.vector_template.shape)
    np.testing.assert_array_equal((v + w).get_values(), a + b)
    np.testing.assert_array_equal((v - w).get_values(), a - b)
    np.testing.assert_array_equal((v * 1.7).get_values(), a * 1.7)
    np.testing.assert_array_equal((1.7 * v).get_values(), a * 1.7)
    assert abs(w.norm() - np.linalg.norm(b)) <= 1e-14 * np.linalg.norm(b)
    z = v.clone_zero()
    assert not np.any(z.get_values())
    c = v.clone()
    c.set_values(a * 2)
    np.testing.assert_array_equal(v.get_values(), a)                 # clone does not alias
    u = v.clone_zero()
    u.unpack(c.pack())
    np.testing.assert_array_equal(u.get_values(), a * 2)
    v2 = copy.deepcopy(v)
    np.testing.assert_array_equal(v2.get_values(), a)


def test_solver_state_is_user_visible(P):
    """output_fcn / user code reads self.u[0][i], self.t[0], self.index_local_c[0] (docs/source/usage/advanced.rst)."""
    prob = P.simple_setup_problem(P.Heat2D(x_start=0, x_end=1, y_start=0, y_end=1, nx=17, ny=21, a=1, rhs=C.heat2d_rhs,
                                           t_start=0, t_stop=1, nt=17), level=2, coarsening=4)
    seen = {}

    def output(self):
        seen['last'] = self.u[0][-1].get_values().copy()
        seen['t'] = self.t[0][-1]
        seen['c'] = list(self.index_local_c[0])

    import logging
    solver = P.Mgrit(problem=prob, output_fcn=output, tol=1e-9, logging_lvl=logging.WARNING)
    solver.solve()
    assert seen['last'].shape == (17, 21) and seen['t'] == 1.0 and seen['c'][:2] == [0, 4]
    assert np.array_equal(seen['last'][0], np.zeros(21))              # Dirichlet boundary
    # writing a vector back and reading it again round-trips through the level storage
    v = solver.u[0][3]
    solver.u[0][5] = v
    assert np.max(np.abs(solver.u[0][5].get_values() - v.get_values())) <= 1e-13


@pytest.mark.parametrize('variant', ['uniform', 'nonuniform_t', 'zero_rhs', 'rank2_rhs', 'nx1001_product'])
def test_spectral_coarse_solve_matches_phi_chain(P, variant):
    """The coarsest-level solve in sine space (csrc/spectral.cu) against the chain of tridiagonal Phi applications it
    replaces (mgrit.py:459-486), on the same u[0] and FAS right-hand side g: 1e-12 relative."""
    import logging
    import torch
    kw = dict(x_start=0, x_end=1, nx=1025, a=1, init_cond=C.heat_init, rhs=C.heat_rhs, sine_space=False)
    if variant == 'zero_rhs':
        kw.pop('rhs')
    if variant == 'rank2_rhs':
        kw['rhs'] = C.heat_rhs_rank2
    if variant == 'nx1001_product':          # n + 1 not a power of two: the product with the sine matrix, no FFT
        kw['nx'] = 1001
    t = np.linspace(0, 2, 513)
    if variant == 'nonuniform_t':
        t = 2 * np.linspace(0, 1, 513) ** 1.3
    fine = P.Heat1D(t_interval=t, **kw)
    coarse = P.Heat1D(t_interval=t[::2], **kw)
    solver = P.Mgrit(problem=[fine, coarse], nested_iteration=False, logging_lvl=logging.WARNING)
    assert 1 in solver._spectral
    assert solver._spectral[1].xform.fast == (variant != 'nx1001_product')
    lv = solver._lv[1]
    gen = torch.Generator(device='cuda').manual_seed(7)
    lv.g[:, :lv.n] = torch.randn((lv.npts, lv.n), generator=gen, device='cuda', dtype=torch.float64) * 1e-2
    solver.forward_solve(1)
    spectral = lv.u.clone()
    sp = solver._spectral.pop(1)
    lv.u[1:].zero_()
    solver.forward_solve(1)
    chain = lv.u.clone()
    solver._spectral[1] = sp
    assert float(chain[1:].abs().max()) > 1e-3
    assert float((spectral - chain).abs().max()) <= 1e-12 * float(chain.abs().max())
    assert float(spectral[:, lv.n:].abs().max()) == 0.0           # the row padding stays zero


def test_small_table_uploads_survive_a_busy_device(P):
    """The small Phi tables go to the device through pooled page-locked buffers with asynchronous copies.  While the device
    is busy (here: a queue of matrix products; on several time ranks: the setup of the communicator) the host runs ahead
    of those copies, so a buffer must not be handed out again before its copy has run: every level's natural-order
    table must hold that level's own reciprocals.  (Regression: heat1d_small_f_cf2 stalled at 4.9e-10 in one of eight
    runs on two ranks because a level read another level's 1 / (1 + dt lam).)"""
    import logging
    import torch
    x = torch.randn(6144, 6144, device='cuda', dtype=torch.float64)
    for _ in range(6):                                   # tens of milliseconds of queued device work
        x = (x @ x) * 1e-4
    t = np.linspace(0, 2, 129)
    kw = dict(x_start=0, x_end=1, nx=17, a=1, init_cond=C.heat_init, rhs=C.heat_rhs)
    levels = [P.Heat1D(t_interval=t[::2 ** k], **kw) for k in range(4)]
    solver = P.Mgrit(problem=levels, nested_iteration=False, logging_lvl=logging.WARNING)
    assert solver.problem[0].kind == P._lib.APP_HEAT1D_SINE
    torch.cuda.synchronize()
    for lv in solver._lv:
        host, _ = lv.app.sine_host_tables(lv.t, lv.team_threads, lv.chunk)
        assert np.array_equal(lv.nat_dev[:2].cpu().numpy(), host['nat'])
        for name, ten in (('sconst', lv.c.sconst_dev), ):
            assert ten
    info = solver.solve()
    assert info['conv'][-1] < 1e-7


@pytest.mark.parametrize('n', [1, 2, 3, 5, 16, 20, 33, 100, 300, 1000, 1500, 4095, 4096])
def test_rows_rfft_matches_numpy(P, n):
    """csrc/fourier.cu: the real Fourier transform of rows of any length n <= 4096 (Bluestein's algorithm on a radix-2 FFT
    in shared memory) and its inverse, against numpy.fft: 1e-13 of the largest coefficient."""
    import torch
    from pymgrit_b200.advection.advection_1d import circ_fft_tables
    lib = P._lib.lib()
    dev = torch.device('cuda', torch.cuda.current_device())
    tw, chirp, bhat = circ_fft_tables(n, dev)
    rows, k = 7 + (n & 1), n // 2 + 1                   # rows are transformed in pairs: an odd and an even count
    rng = np.random.default_rng(n)
    a = rng.standard_normal((rows, n + 3))
    row0 = rng.standard_normal(n)
    a_dev, row0_dev = torch.as_tensor(a).to(dev), torch.as_tensor(row0).to(dev)
    c = torch.zeros((rows, 2 * k + 2), dtype=torch.float64, device=dev)
    st = P._lib.current_stream_ptr()
    P._lib.check(lib.mgb_rows_rfft(rows, n, a_dev.data_ptr(), n + 3, row0_dev.data_ptr(), tw.data_ptr(), chirp.data_ptr(),
                                   bhat.data_ptr(), c.data_ptr(), 2 * k + 2, st), 'rows_rfft')
    want = np.fft.rfft(np.vstack([row0[None], a[1:, :n]]), axis=1)
    got = c.cpu().numpy()[:, :2 * k].reshape(rows, k, 2)
    got = got[..., 0] + 1j * got[..., 1]
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want)) * max(1.0, np.log2(n + 1))
    assert float(c[:, 2 * k:].abs().max()) == 0.0
    out = torch.full((rows, n + 1), 7.0, dtype=torch.float64, device=dev)
    P._lib.check(lib.mgb_rows_irfft(rows, 1, n, c.data_ptr(), 2 * k + 2, tw.data_ptr(), chirp.data_ptr(), bhat.data_ptr(),
                                    out.data_ptr(), n + 1, st), 'rows_irfft')
    back = out.cpu().numpy()
    assert np.all(back[0] == 7.0) and np.all(back[:, n] == 7.0)          # row 0 and the padding are not touched
    assert np.max(np.abs(back[1:, :n] - a[1:, :n])) <= 1e-13 * np.max(np.abs(a)) * max(1.0, np.log2(n + 1))


@pytest.mark.parametrize('variant', ['uniform', 'nonuniform_t', 'nx4096', 'nx100_even'])
def test_fourier_coarse_solve_matches_phi_chain(P, variant, monkeypatch):
    """The coarsest-level solve of Advection1D in Fourier space (csrc/fourier.cu: real FFT of all rows, time-parallel
    complex recurrences, inverse FFT) against the chain of cyclic bidiagonal Phi applications it replaces
    (mgrit.py:459-486, advection_1d.py:129-143), on the same u[0] and FAS right-hand side g: 1e-12 relative."""
    import logging
    import torch
    nx = {'nx4096': 4096, 'nx100_even': 101}.get(variant, 258)
    nt = 1025 if variant != 'nx4096' else 385
    t = np.linspace(0, 2, nt)
    if variant == 'nonuniform_t':
        t = 2 * np.linspace(0, 1, nt) ** 1.3
    kw = dict(c=1, x_start=-1, x_end=1, nx=nx)
    fine = P.Advection1D(t_interval=t, **kw)
    coarse = P.Advection1D(t_interval=t[::2], **kw)
    solver = P.Mgrit(problem=[fine, coarse], nested_iteration=False, logging_lvl=logging.WARNING)
    assert type(solver._spectral.get(1)).__name__ == 'FourierSolve'
    lv = solver._lv[1]
    gen = torch.Generator(device='cuda').manual_seed(11)
    lv.g[:, :lv.n] = torch.randn((lv.npts, lv.n), generator=gen, device='cuda', dtype=torch.float64) * 1e-2
    solver.forward_solve(1)
    spectral = lv.u.clone()
    sp = solver._spectral.pop(1)
    lv.u[1:].zero_()
    solver.forward_solve(1)
    chain = lv.u.clone()
    solver._spectral[1] = sp
    assert float(chain[1:].abs().max()) > 1e-3
    assert torch.equal(spectral[0], chain[0])                            # the start value is not rewritten
    assert float((spectral - chain)[:, :lv.n].abs().max()) <= 1e-12 * float(chain[:, :lv.n].abs().max())
    if lv.pitch > lv.n:
        assert float(spectral[:, lv.n:].abs().max()) == 0.0              # the row padding stays zero
    # and the whole solve: same iteration count and history with and without it
    info_f = solver.solve()
    monkeypatch.setenv('MGB_ADVECTION_FOURIER', '0')
    plain = P.Mgrit(problem=[P.Advection1D(t_interval=t, **kw), P.Advection1D(t_interval=t[::2], **kw)],
                    nested_iteration=False, logging_lvl=logging.WARNING)
    assert not plain._spectral
    info_c = plain.solve()
    assert len(info_f['conv']) == len(info_c['conv'])
    assert np.allclose(info_f['conv'], info_c['conv'], rtol=1e-6, atol=1e-13)
    n0 = solver._lv[0].n
    uf, uc = solver._lv[0].u[:, :n0], plain._lv[0].u[:, :n0]
    assert float((uf - uc).abs().max()) <= 1e-11 * float(uc.abs().max())


@pytest.mark.parametrize('variant', ['uniform', 'nonuniform_t', 'zero_rhs', 'rank2_rhs', 'nx1001_product', 'short'])
def test_sine_level_solve_matches_phi_chain(P, variant):
    """A level kept in sine space: the time-parallel scalar recurrences of mgb_sine_level_solve against (a) the chain of
    diagonal Phi applications through the generic sweep kernel and (b) the chain of tridiagonal solves on the same data in
    node space (mgrit.py:459-486): 1e-12 relative."""
    import logging
    import torch
    kw = dict(x_start=0, x_end=1, nx=1025, a=1, init_cond=C.heat_init, rhs=C.heat_rhs)
    if variant == 'zero_rhs':
        kw.pop('rhs')
    if variant == 'rank2_rhs':
        kw['rhs'] = C.heat_rhs_rank2
    if variant == 'nx1001_product':
        kw['nx'] = 1001
    nt = 37 if variant == 'short' else 513
    t = np.linspace(0, 2, nt)
    if variant == 'nonuniform_t':
        t = 2 * np.linspace(0, 1, nt) ** 1.3
    out = {}
    g_nodes = None
    for space in ('sine', 'node'):
        fine = P.Heat1D(t_interval=t, sine_space=space == 'sine', **kw)
        coarse = P.Heat1D(t_interval=t[::2], sine_space=space == 'sine', **kw)
        solver = P.Mgrit(problem=[fine, coarse], nested_iteration=False, logging_lvl=logging.WARNING)
        assert solver.problem[1].kind == (P._lib.APP_HEAT1D_SINE if space == 'sine' else P._lib.APP_HEAT1D)
        lv = solver._lv[1]
        if g_nodes is None:
            gen = torch.Generator(device='cuda').manual_seed(7)
            g_nodes = torch.randn((lv.npts, lv.n), generator=gen, device='cuda', dtype=torch.float64) * 1e-2
        lv.app.values_to_rows(g_nodes, lv.g)                       # the same FAS right-hand side in either representation
        sp = solver._spectral.pop(1, None)
        solver.forward_solve(1)                                    # chain of Phi applications (mgb_forward_solve)
        out[space, 'chain'] = lv.app.rows_to_values(lv.u).clone()
        if space == 'sine':
            assert type(sp).__name__ == 'SineLevelSolve'
            lv.u[1:].zero_()
            solver._spectral[1] = sp
            solver.forward_solve(1)
            out[space, 'solve'] = lv.app.rows_to_values(lv.u).clone()
            assert float(lv.u[:, lv.n:].abs().max()) == 0.0        # the row padding stays zero
    ref = out['node', 'chain']
    scale = float(ref.abs().max())
    assert scale > 1e-3
    assert float((out['sine', 'chain'] - ref).abs().max()) <= 1e-12 * scale
    assert float((out['sine', 'solve'] - ref).abs().max()) <= 1e-12 * scale


@pytest.mark.parametrize('name', ['heat1d_cfg2_nt1025', 'heat1d_nonuniform_t', 'heat1d_varying', 'heat1d_small_f_cf2',
                                  'heat1d_rhs_rank2', 'heat1d_example'])
def test_sine_space_solve_matches_node_space_solve(P, name, monkeypatch):
    """The whole MGRIT solve with the level rows in sine space against the tridiagonal kernels: same iteration count,
    residual history to 1e-12 of the run's scale, solution to 1e-12 relative."""
    from b200_util import run_b200, solution_rows
    if name not in C.CASES:
        pytest.skip('case not defined')
    monkeypatch.setenv('MGB_HEAT1D_SINE', '1')
    sine, info_s = run_b200(name)
    assert sine.problem[0].kind == P._lib.APP_HEAT1D_SINE
    monkeypatch.setenv('MGB_HEAT1D_SINE', '0')
    node, info_n = run_b200(name)
    assert node.problem[0].kind == P._lib.APP_HEAT1D
    us, un = solution_rows(sine)[0], solution_rows(node)[0]
    scale = np.max(np.abs(un))
    assert len(info_s['conv']) == len(info_n['conv'])
    assert np.max(np.abs(info_s['conv'] - info_n['conv'])) <= 1e-12 * scale * np.sqrt(len(un))
    assert np.max(np.abs(us - un)) <= 1e-12 * scale


@pytest.mark.parametrize('name', ['heat1d_cfg2_nt1025', 'heat1d_rhs_nonsep', 'heat1d_trailing_f', 'advection_cfg4_small',
                                  'heat2d_cfg3_small', 'heat1d_bdf2_example', 'heat1d_weighted'])
def test_lazy_f_points_are_bit_identical(P, name, monkeypatch):
    """Level 0 stores only the last F-point of every interval while the solver iterates and computes the others once at
    the end (MGB_CORRECT_LAST_ONLY + one F-relaxation): same residual history and solution, bit for bit, as storing
    every F-point in every iteration."""
    from b200_util import run_b200, solution_rows
    monkeypatch.setenv('MGB_LAZY_F', '1')
    lazy, info_l = run_b200(name)
    monkeypatch.setenv('MGB_LAZY_F', '0')
    eager, info_e = run_b200(name)
    assert np.array_equal(info_l['conv'], info_e['conv'])
    assert np.array_equal(solution_rows(lazy)[0], solution_rows(eager)[0])


@pytest.mark.parametrize('name', ['heat1d_cfg2_nt1025', 'heat1d_nonuniform_t', 'heat1d_rhs_nonsep', 'heat1d_small_f_cf2',
                                  'heat1d_rhs_rank2', 'heat1d_trailing_f', 'heat1d_nx4097', 'advection_cfg4_small',
                                  'heat2d_bc', 'heat2d_cfg3_small', 'heat1d_bdf2_example', 'heat1d_bdf2_small_f'])
def test_fused_down_sweep_is_bit_identical(P, name, monkeypatch):
    """mgb_down_sweep (C-relaxation + F-relaxation + FAS restriction in one launch) against the three separate
    launches: same residual history and the same solution, bit for bit."""
    from b200_util import run_b200, solution_rows
    monkeypatch.setenv('MGB_FUSED_DOWN', '1')
    fused, info_f = run_b200(name)
    assert any(fused._fused_down)
    monkeypatch.setenv('MGB_FUSED_DOWN', '0')
    plain, info_p = run_b200(name)
    assert not any(plain._fused_down)
    assert np.array_equal(info_f['conv'], info_p['conv']), (info_f['conv'], info_p['conv'])
    assert np.array_equal(solution_rows(fused)[0], solution_rows(plain)[0])


@pytest.mark.parametrize('method', ['CN', 'FE', 'BE'])
def test_heat2d_theta_method_step_against_oracle(P, method):
    """Heat2D.step for the three branches of heat_2d.py:322-366 against the oracle's sparse solve of the same system."""
    from oracle import mgrit_oracle as O
    kw = dict(x_start=0, x_end=1, y_start=0, y_end=2, nx=33, ny=21, a=0.7, rhs=C.heat2d_rhs, init_cond=C.heat2d_init,
              method=method)
    if method != 'FE':
        kw.update(bc_left=1.0, bc_bottom=lambda y: 0.5 * y)
    app = P.Heat2D(t_start=0, t_stop=1, nt=11, **kw)
    orc = O.Heat2DOracle(t_start=0, t_stop=1, nt=11, **kw)
    dt = 1e-4 if method == 'FE' else 0.1
    got = app.step(u_start=app.vector_t_start, t_start=0.2, t_stop=0.2 + dt).get_values()
    ref = orc.phi(orc.u0, 0.2, 0.2 + dt)
    assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref))
    if method == 'FE':
        with pytest.raises(Exception):
            P.Heat2D(t_start=0, t_stop=1, nt=11, bc_left=1.0, **kw)


def test_allen_cahn_known_answers_and_contract(P):
    """tests/allen_cahn/test_allen_cahn.py: constructor, the initial condition, the IMEX step (test_heat_2d_step_imex),
    the Vector arithmetic; the Newton branches raise (not built)."""
    kw = dict(nx=3, eps=3, newton_tol=5, newton_maxiter=4, lin_tol=5, lin_maxiter=1, radius=0.25, nu=1, t_start=0, t_stop=1,
              nt=11)
    app = P.AllenCahn(method='IMEX', **kw)
    assert (app.nx, app.nu, app.eps, app.newton_maxiter, app.newton_tol, app.lin_tol, app.lin_maxiter, app.radius,
            app.method) == (3, 1, 3, 4, 5, 5, 1, 0.25, 'IMEX')
    assert isinstance(app.vector_template, P.VectorAllenCahn2D) and isinstance(app.vector_t_start, P.VectorAllenCahn2D)
    np.testing.assert_almost_equal(app.vector_t_start.get_values(), np.array([[-0.10732614, -0.05885746, -0.10732614],
                                                                              [-0.05885746, 0.05885746, -0.05885746],
                                                                              [-0.10732614, -0.05885746, -0.10732614]]))
    res = app.step(u_start=app.vector_t_start, t_start=0, t_stop=0.1)
    np.testing.assert_almost_equal(res.get_values(), np.array([[-0.07997795, -0.0640509, -0.07997795],
                                                               [-0.0640509, -0.03719789, -0.0640509],
                                                               [-0.07997795, -0.0640509, -0.07997795]]))
    with pytest.raises(Exception):
        P.AllenCahn(method='DE', **kw)
    with pytest.raises(Exception):
        P.AllenCahn(method='IMPL', **kw)                          # Newton branches: no device kernels
    v1, v2 = P.VectorAllenCahn2D(3, 3), P.VectorAllenCahn2D(3, 3)
    v1.set_values(np.ones((3, 3)))
    v2.set_values(2 * np.ones((3, 3)))
    np.testing.assert_equal((v1 + v2).get_values(), 3 * np.ones((3, 3)))
    np.testing.assert_equal((v2 - v1).get_values(), np.ones((3, 3)))
    np.testing.assert_equal((v1 * 5).get_values(), 5 * np.ones((3, 3)))
    v1.set_values(np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9.]]))
    assert abs(v1.norm() - np.linalg.norm(np.arange(1, 10.))) < 1e-14


def test_allen_cahn_step_against_oracle_at_example_size(P):
    """One IMEX step at the example's size (128 x 128 periodic nodes) against the oracle's sparse direct solve."""
    from oracle import mgrit_oracle as O
    app = P.AllenCahn(t_start=0, t_stop=0.032, nt=33, method='IMEX')
    orc = O.AllenCahnOracle(t_start=0, t_stop=0.032, nt=33, method='IMEX')
    got = app.step(u_start=app.vector_t_start, t_start=0.0, t_stop=0.001).get_values()
    ref = orc.phi(orc.u0, 0.0, 0.001)
    assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_user_defined_batched_application(P):
    """The extension point (core/batched.py): a user application that only implements step_rows
    with its own device code -- here torch operations for the Dahlquist test equation -- runs through Mgrit and gives the
    reference's residual history (README.rst:102-122)."""
    import torch

    class MyVector(P.DeviceVector):
        def __init__(self, tensor=None):
            super().__init__((1,), tensor)

    class MyDahlquist(P.BatchedApplication):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            self.ndof = 1
            self.vector_template = MyVector()
            self.vector_t_start = MyVector()
            self.vector_t_start.set_values(np.array([1.0]))

        def step_rows(self, src, src_idx, dst, dst_idx, t_start, t_stop):
            dt = torch.as_tensor(t_stop - t_start, device=src.device)
            vals = src[src_idx.long(), 0] / (1 + dt)                       # backward Euler, lambda = -1
            dst[dst_idx.long(), 0] = vals

    prob = P.simple_setup_problem(MyDahlquist(t_start=0, t_stop=5, nt=101), level=2, coarsening=2)
    import logging
    info = P.Mgrit(problem=prob, tol=1e-10, logging_lvl=logging.WARNING).solve()
    ref = [7.186185937025427e-05, 1.246106707585954e-06, 2.1015566149418615e-08, 3.1441273895579124e-10, 3.975216519949153e-12]
    assert len(info['conv']) == 5 and np.max(np.abs(np.array(info['conv']) - ref)) <= 1e-14


@pytest.mark.parametrize('name', ['example_dahlquist', 'example_heat_1d', 'example_heat_1d_bdf2', 'example_spatial_coarsening',
                                  'example_heat_2d', 'example_at_mgrit', 'example_allen_cahn'])
def test_examples_run(name):
    """examples/*.py -- the reference's examples with the import swapped -- run to the end on the device and report a
    residual history that went down."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get('PYTHONPATH', ''))
    r = subprocess.run([sys.executable, os.path.join(root, 'examples', name + '.py')], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600, env=env, cwd=os.path.join(root, 'examples'))
    assert r.returncode == 0, r.stdout[-3000:]
    assert 'conv' in r.stdout or re.search(r'\[[0-9.e+\- \n]+\]', r.stdout), r.stdout[-2000:]


@pytest.mark.parametrize('name', ['heat1d_cfg2_nt1025', 'heat1d_nonuniform_t', 'heat1d_rhs_rank2', 'heat1d_trailing_f',
                                  'heat1d_small_f_cf2', 'heat1d_example', 'heat1d_zero_rhs', 'heat1d_nx4097'])
def test_one_thread_per_mode_sweeps_are_bit_identical_to_the_team_kernels(P, name, monkeypatch):
    """csrc/sine_modes.cu (one thread per mode, no shared-memory staging) against the team kernels of csrc/sweeps.cuh on
    sine-space levels: the same operations per element in the same order, so the same bits -- residual history (whose
    per-row sums of squares are added in another order: 1e-15 relative) and solution."""
    from b200_util import run_b200, solution_rows
    if name not in C.CASES:
        pytest.skip('case not defined')
    monkeypatch.setenv('MGB_HEAT1D_SINE', '1')
    monkeypatch.setenv('MGB_SINE_MODES', '1')
    modes, info_m = run_b200(name)
    assert modes.problem[0].kind == P._lib.APP_HEAT1D_SINE
    monkeypatch.setenv('MGB_SINE_MODES', '0')
    team, info_t = run_b200(name)
    assert len(info_m['conv']) == len(info_t['conv'])
    assert np.allclose(info_m['conv'], info_t['conv'], rtol=1e-12, atol=0)
    assert np.array_equal(solution_rows(modes)[0], solution_rows(team)[0])


@pytest.mark.parametrize('name', ['heat2d_example', 'heat2d_cfg3_small', 'heat2d_bc', 'heat2d_cn', 'heat2d_cn_3lvl',
                                  'heat2d_fe'])
def test_heat2d_one_thread_per_mode_sweeps_are_bit_identical_to_the_team_kernels(P, name, monkeypatch):
    """Heat2D rows always hold sine coefficients, so the hot sweeps run as one thread per coefficient too
    (csrc/sine_modes.cu TilePhi: backward Euler on uniform levels, and the general theta / several-step-size variant);
    same operations per element as the tile teams of csrc/sweeps.cuh, Dirichlet tiles included."""
    from b200_util import run_b200, solution_rows
    if name not in C.CASES:
        pytest.skip('case not defined')
    monkeypatch.setenv('MGB_SINE_MODES', '1')
    modes, info_m = run_b200(name)
    monkeypatch.setenv('MGB_SINE_MODES', '0')
    team, info_t = run_b200(name)
    assert len(info_m['conv']) == len(info_t['conv'])
    assert np.allclose(info_m['conv'], info_t['conv'], rtol=1e-10, atol=1e-300)
    assert np.array_equal(solution_rows(modes)[0], solution_rows(team)[0])
