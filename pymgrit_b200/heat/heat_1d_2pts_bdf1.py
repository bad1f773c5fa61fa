"""1-D heat equation, pairs of time points advanced by two backward-Euler (BDF1) steps, with the reference's interface
(heat/heat_1d_2pts_bdf1.py:16-117); state and Phi on the GPU (pymgrit_b200/heat/heat_1d_2pts.py)."""
from pymgrit_b200.core.vector import DeviceVector
from pymgrit_b200.heat.heat_1d_2pts import Heat1D2Pts


class Heat1DBDF1(Heat1D2Pts):
    method = 'BDF1'

    def _second_start_value(self, first):
        # one BDF1 step from t_0 to t_0 + dtau (heat_1d_2pts_bdf1.py:64-66) = Heat1D's device step
        t0 = float(self.t[0])
        app = self._heat1d(t0, t0 + self.dtau)
        u = app.vector_template.clone_zero()
        u.set_values(first)
        return app.step(u_start=u, t_start=t0, t_stop=t0 + self.dtau).device_values

    def _coefficients(self, dt):
        # heat_1d_2pts_bdf1.py:108-113: a step over (dt - dtau) into t_stop, then one over dtau
        fac = self.a / self.dx ** 2
        tau = dt - self.dtau
        return tau * fac, self.dtau * fac, 0.0, 1.0, 0.0, 1.0, tau, self.dtau
