"""The closed form behind `LagCoef` / `rk4_step<.., LAG>` (csrc/d2dx_device.cuh): for phi' = -(phi - phi_c) / tau with the
input held, every classical RK4 stage value and the step itself are phi_c + (phi - phi_c) * c_i(z), z = -h / tau.  This
checks the factor recurrences the kernel uses against the four stage derivatives written out (d2d/dynamic.py:14-23 has the
lag, the oracle's rk4 the stages), over stiff and slack time constants -- including z = -5, outside RK4's stability interval,
where the formation scripts' dt = 0.05 with tau_phi = 0.01 would sit without sub-steps."""
import numpy as np


def lag_coef(h, n_inv_tau):
    z = h * n_inv_tau
    c2 = 0.5 * z + 1.0
    c3 = 0.5 * z * c2 + 1.0
    c4 = z * c3 + 1.0
    cf = z / 6.0 * ((1.0 + 2.0 * c2) + (2.0 * c3 + c4)) + 1.0
    return c2, c3, c4, cf


def rk4_lag(phi, phi_c, h, n_inv_tau):
    f = lambda p: n_inv_tau * (p - phi_c)
    k1 = f(phi); p2 = phi + 0.5 * h * k1
    k2 = f(p2); p3 = phi + 0.5 * h * k2
    k3 = f(p3); p4 = phi + h * k3
    k4 = f(p4)
    return p2, p3, p4, phi + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


def test_stage_factors_reproduce_the_rk4_stages_of_a_first_order_lag():
    rng = np.random.default_rng(0)
    for tau in (0.01, 0.05, 0.3, 1.0, 7.0):
        for h in (0.002, 0.01, 0.05):
            phi, phi_c = rng.normal(0, 0.5, 1000), rng.normal(0, 0.5, 1000)
            c = lag_coef(h, -1.0 / tau)
            ref = rk4_lag(phi, phi_c, h, -1.0 / tau)
            scale = max(1.0, abs(h / tau)) ** 4
            for ci, ri in zip(c, ref):
                np.testing.assert_allclose(phi_c + (phi - phi_c) * ci, ri, rtol=0, atol=2e-15 * scale)
    # the step factor is RK4's stability polynomial 1 + z + z^2/2 + z^3/6 + z^4/24
    for z in (-5.0, -1.0, -0.2, -1e-3):
        assert abs(lag_coef(1.0, z)[3] - (1 + z + z * z / 2 + z ** 3 / 6 + z ** 4 / 24)) < 1e-14 * max(1.0, abs(z)) ** 4
