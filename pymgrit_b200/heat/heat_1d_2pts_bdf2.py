"""1-D heat equation, pairs of time points advanced by two variable-step BDF2 steps, with the reference's interface
(heat/heat_1d_2pts_bdf2.py:17-138); state and Phi on the GPU (pymgrit_b200/heat/heat_1d_2pts.py)."""
import numpy as np

from pymgrit_b200.heat.heat_1d_2pts import Heat1D2Pts


class Heat1DBDF2(Heat1D2Pts):
    method = 'BDF2'

    def _second_start_value(self, first):
        # trapezoidal rule (heat_1d_2pts_bdf2.py:65-68):
        #   (I + dtau/2 L) u1 = (I - dtau/2 L) u0 + dtau/2 (b(t0) + b(t0 + dtau))
        # = Heat1D's backward-Euler device step over dtau/2 ending at t0 + dtau, applied to
        #   (I - dtau/2 L) u0 + dtau/2 b(t0)   (a three-point stencil and one evaluation of b on the host)
        t0 = float(self.t[0])
        half = self.dtau / 2
        fac = self.a / self.dx ** 2
        lu = 2 * fac * first
        lu[1:] -= fac * first[:-1]
        lu[:-1] -= fac * first[1:]
        start = first - half * lu + half * np.asarray(self.rhs(self.x, t0), dtype=float)
        app = self._heat1d(t0 + half, t0 + self.dtau)
        u = app.vector_template.clone_zero()
        u.set_values(start)
        return app.step(u_start=u, t_start=t0 + half, t_stop=t0 + self.dtau).device_values

    def _coefficients(self, dt):
        # heat_1d_2pts_bdf2.py:110-133, each system (L + coeff I) y = b - coeffm2 u_{-2} + coeffm1 u_{-1} divided by coeff
        fac = self.a / self.dx ** 2
        dtau = self.dtau
        tau_i, tau_im1 = dt - dtau, dtau
        r_i = tau_i / tau_im1
        cm2 = (r_i ** 2) / (tau_i * (1 + r_i))
        cm1 = (1 + r_i) / tau_i
        co = (1 + 2 * r_i) / (tau_i * (1 + r_i))
        r1, a1, b1, c1 = fac / co, -cm2 / co, cm1 / co, 1.0 / co
        tau_im1, tau_i = tau_i, dtau
        r_i = tau_i / tau_im1
        cm2 = (r_i ** 2) / (tau_i * (1 + r_i))
        cm1 = (1 + r_i) / tau_i
        co = (1 + 2 * r_i) / (tau_i * (1 + r_i))
        r2, a2, b2, c2 = fac / co, -cm2 / co, cm1 / co, 1.0 / co
        return r1, r2, a1, b1, a2, b2, c1, c2
