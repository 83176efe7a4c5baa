"""CPU tier: the torchdiffeq restatement (oracle/ode_oracle.py) cross-checked against SciPy's RK45 -
same Dormand-Prince tableau and 0.9/0.2/10 controller (the solver itself has no
reference-side fixture, so this is its independent anchor)."""
import numpy as np
import torch
from scipy.integrate import solve_ivp

from oracle import ode_oracle


def _lotka(t, y):
    a, b, c, d = 1.5, 1.0, 3.0, 1.0
    return np.stack([a * y[0] - b * y[0] * y[1], -c * y[1] + d * y[0] * y[1]])


def test_dopri5_matches_scipy_rk45():
    y0 = np.array([10.0, 5.0])
    t_eval = np.linspace(0.0, 2.0, 9)
    ref = solve_ivp(_lotka, (0.0, 2.0), y0, method="RK45", rtol=1e-7, atol=1e-9, t_eval=t_eval)
    stats = {}

    def f(t, y):
        return torch.from_numpy(_lotka(float(t), y.numpy()))

    sol = ode_oracle.odeint(f, torch.from_numpy(y0), torch.from_numpy(t_eval), rtol=1e-7, atol=1e-9,
                            method="dopri5", stats=stats)
    np.testing.assert_allclose(sol.numpy().T, ref.y, rtol=2e-6)
    # same controller => comparable work (scipy clips the last step, torchdiffeq overshoots)
    assert abs(stats["nfe"] - ref.nfev) <= 0.2 * ref.nfev


def test_fixed_grid_orders():
    """euler / midpoint / rk4(3/8) converge at orders 1 / 2 / 4 on y' = -y."""
    def f(t, y):
        return -y
    y0 = torch.tensor([1.0], dtype=torch.float64)
    errs = {}
    for method in ("euler", "midpoint", "rk4"):
        e = []
        for n in (11, 21):
            t = torch.linspace(0, 1, n, dtype=torch.float64)
            e.append(abs(float(ode_oracle.odeint(f, y0, t, method=method)[-1]) - np.exp(-1.0)))
        errs[method] = np.log2(e[0] / e[1])
    assert 0.9 < errs["euler"] < 1.1 and 1.9 < errs["midpoint"] < 2.1 and 3.8 < errs["rk4"] < 4.2


def test_tuple_state_and_reverse_time():
    def f(t, state):
        a, b = state
        return (-a, 2.0 * b)
    y0 = (torch.ones(3, 2), torch.full((4,), 0.5))
    t = torch.linspace(0.0, 1.0, 5)
    a, b = ode_oracle.odeint(f, y0, t, rtol=[1e-6, 1e-6], atol=[1e-8, 1e-8], method="dopri5")
    assert a.shape == (5, 3, 2) and b.shape == (5, 4)
    np.testing.assert_allclose(a[-1].numpy(), np.exp(-1.0), rtol=1e-4)
    np.testing.assert_allclose(b[-1].numpy(), 0.5 * np.exp(2.0), rtol=1e-4)
    # decreasing grid: integrate back to the start
    back = ode_oracle.odeint(lambda t, y: -y, a[-1], torch.linspace(1.0, 0.0, 5), rtol=1e-6, atol=1e-8, method="dopri5")
    np.testing.assert_allclose(back[-1].numpy(), 1.0, rtol=1e-4)
