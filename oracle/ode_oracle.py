"""TEST INFRASTRUCTURE ONLY (oracle/).  CPU restatement of the subset of torchdiffeq==0.2.5
(ti_env.yml:15 - a pip pin, NOT vendored under /root/reference) that the reference's
samplers reach:

    mdqm9/thermo/ambient/integrators.py:45-53,58-66   odeint_adjoint(func, y0, t, method, atol, rtol, adjoint_params)
    mdqm9/thermo/latent/integrators.py:66-74,79-87
    adw/thermo/integrators.py:49-55,60-66

PARITY STATUS: **unpinned** for the solver arithmetic itself - torchdiffeq is absent from this
image and from /opt/wheelhouse, and the reference holds no tests or golden vectors for it.  The
algorithm below follows torchdiffeq's published `_impl/{odeint,adjoint,misc,rk_common,dopri5,
fixed_grid,interp,solvers}.py`; tests cross-validate dopri5 against
`scipy.integrate.solve_ivp(method="RK45")` (same tableau / controller constants).

What is mirrored:
  * forward solve under torch.no_grad() (the adjoint machinery is never used by the samplers)
  * tuple states are flattened with reshape(-1)+cat and un-flattened by the y0 shapes;
    per-component tolerances expand to per-element vectors
  * `t` is cast to the state dtype before every call of `func`
  * decreasing time grids are solved as increasing ones on (-t, -f)
  * fixed grid: euler / midpoint / rk4 (3/8 rule) on the output grid itself, time in the grid dtype
  * dopri5: time-like quantities in float64, state in its own dtype, stage times computed in
    the state dtype, alpha==1 stages evaluated at nextafter(t1, -inf), FSAL, error ratio =
    norm(err / (atol + rtol*max|y0|,|y1|)) with RMS norm (tensor state) or max of per-component
    RMS (tuple state), controller 0.9 / 0.2 / 10 with exponent 1/5, Hairer initial step,
    no clipping to the end time, quartic dense output through (y0, y_mid, y1, f0, f1).
"""
from __future__ import annotations

import torch

# ----------------------------------------------------------------------------------------------
# Dormand-Prince 5(4) tableau  (torchdiffeq _impl/dopri5.py)
# ----------------------------------------------------------------------------------------------
DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
DP_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
DP_C_ERROR = [
    35 / 384 - 1951 / 21600,
    0,
    500 / 1113 - 22642 / 50085,
    125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400,
    11 / 84 - 649 / 6300,
    -1.0 / 60.0,
]
DP_C_MID = [
    6025192743 / 30085553152 / 2,
    0,
    51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2,
    11237099 / 235043384 / 2,
]


# ----------------------------------------------------------------------------------------------
# misc helpers (torchdiffeq _impl/misc.py)
# ----------------------------------------------------------------------------------------------
def _rms_norm(tensor):
    return tensor.abs().pow(2).mean().sqrt()


def _flat_to_shape(tensor, length, shapes):
    out, total = [], 0
    for shape in shapes:
        n = 1
        for s in shape:
            n *= s
        out.append(tensor[..., total:total + n].view((*length, *shape)))
        total += n
    return tuple(out)


def _tuple_tol(tol, shapes):
    if isinstance(tol, (list, tuple)):
        assert len(tol) == len(shapes)
        parts = []
        for t, shape in zip(tol, shapes):
            n = 1
            for s in shape:
                n *= s
            # torch.as_tensor(python float) is float32 here: the tolerance is rounded to fp32 before
            # the solver widens it to float64 (misc._tuple_tol) - mirrored on purpose.
            parts.append(torch.as_tensor(float(t)).expand(n))
        return torch.cat(parts)
    return tol


def _nextafter_prev(t):
    return torch.nextafter(t, t - 1)


def _nextafter_next(t):
    return torch.nextafter(t, t + 1)


class _Func:
    """`_TupleFunc` + `_ReverseFunc` + `_PerturbFunc` rolled into one callable."""

    def __init__(self, func, shapes, reverse):
        self.func, self.shapes, self.reverse = func, shapes, reverse
        self.nfe = 0

    def __call__(self, t, y, perturb=0):
        t = t.to(y.abs().dtype)  # "this dtype change here might be buggy" - mirrored on purpose
        if perturb > 0:
            t = _nextafter_next(t)
        elif perturb < 0:
            t = _nextafter_prev(t)
        if self.reverse:
            t = -t
        self.nfe += 1
        if self.shapes is not None:
            f = self.func(t, _flat_to_shape(y, (), self.shapes))
            f = torch.cat([f_.reshape(-1) for f_ in f])
        else:
            f = self.func(t, y)
        return -f if self.reverse else f


# ----------------------------------------------------------------------------------------------
# fixed-grid solvers (torchdiffeq _impl/solvers.py FixedGridODESolver, fixed_grid.py)
# ----------------------------------------------------------------------------------------------
_ONE_THIRD = 1 / 3
_TWO_THIRDS = 2 / 3


def _step_euler(func, t0, dt, t1, y0):
    f0 = func(t0, y0)
    return dt * f0


def _step_midpoint(func, t0, dt, t1, y0):
    half_dt = 0.5 * dt
    f0 = func(t0, y0)
    y_mid = y0 + f0 * half_dt
    return dt * func(t0 + half_dt, y_mid)


def _step_rk4(func, t0, dt, t1, y0):
    # torchdiffeq's "rk4" is the 3/8-rule variant (rk_common.rk4_alt_step_func)
    k1 = func(t0, y0)
    k2 = func(t0 + dt * _ONE_THIRD, y0 + dt * k1 * _ONE_THIRD)
    k3 = func(t0 + dt * _TWO_THIRDS, y0 + dt * (k2 - k1 * _ONE_THIRD))
    k4 = func(t1, y0 + dt * (k1 - k2 + k3))
    return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


_FIXED = {"euler": _step_euler, "midpoint": _step_midpoint, "rk4": _step_rk4}


def _integrate_fixed(step, func, y0, t):
    solution = torch.empty(len(t), *y0.shape, dtype=y0.dtype, device=y0.device)
    solution[0] = y0
    for j in range(1, len(t)):
        t0, t1 = t[j - 1], t[j]
        dt = t1 - t0
        y0 = y0 + step(func, t0, dt, t1, y0)
        solution[j] = y0  # output grid == step grid, so no interpolation
    return solution


# ----------------------------------------------------------------------------------------------
# adaptive dopri5 (torchdiffeq _impl/rk_common.py RKAdaptiveStepsizeODESolver)
# ----------------------------------------------------------------------------------------------
def _select_initial_step(func, t0, y0, order, rtol, atol, norm, f0):
    dtype, t_dtype = y0.dtype, t0.dtype
    t0 = t0.to(t_dtype)
    scale = atol + torch.abs(y0) * rtol
    d0 = norm(y0 / scale).abs()
    d1 = norm(f0 / scale).abs()
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=dtype)
    else:
        h0 = 0.01 * d0 / d1
    h0 = h0.abs()
    y1 = y0 + h0 * f0
    f1 = func(t0 + h0, y1)
    d2 = torch.abs(norm((f1 - f0) / scale) / h0)
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=dtype), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    h1 = h1.abs()
    return torch.min(100 * h0, h1).to(t_dtype)


def _optimal_step_size(last_step, error_ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5):
    if error_ratio == 0:
        return last_step * ifactor
    if error_ratio < 1:
        dfactor = 1.0
    error_ratio = error_ratio.type_as(last_step)
    exponent = 1.0 / order
    factor = min(ifactor, max(safety / float(error_ratio) ** exponent, dfactor))
    return last_step * factor


def _interp_fit(y0, y1, y_mid, f0, f1, dt):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    e = y0
    return [e, d, c, b, a]


def _interp_evaluate(coefficients, t0, t1, t):
    assert (t0 <= t) & (t <= t1), f"invalid interpolation, fails t0 <= t <= t1: {t0}, {t}, {t1}"
    x = ((t - t0) / (t1 - t0)).to(coefficients[0].dtype)
    total = coefficients[0] + x * coefficients[1]
    x_power = x
    for coefficient in coefficients[2:]:
        x_power = x_power * x
        total = total + x_power * coefficient
    return total


def _integrate_dopri5(func, y0, t, rtol, atol, norm, stats):
    sdtype = y0.dtype
    tdtype = torch.promote_types(torch.float64, sdtype)
    rtol = torch.as_tensor(rtol, dtype=tdtype)
    atol = torch.as_tensor(atol, dtype=tdtype)
    alpha = torch.tensor(DP_ALPHA, dtype=torch.float64).to(sdtype)
    beta = [torch.tensor(b, dtype=torch.float64).to(sdtype) for b in DP_BETA]
    c_error = torch.tensor(DP_C_ERROR, dtype=torch.float64).to(sdtype)
    c_mid = torch.tensor(DP_C_MID, dtype=torch.float64).to(sdtype)

    solution = torch.empty(len(t), *y0.shape, dtype=sdtype)
    solution[0] = y0
    t = t.to(tdtype)

    f0 = func(t[0], y0)
    dt = _select_initial_step(func, t[0], y0, 4, rtol, atol, norm, f0)
    t0 = t1 = t[0]
    y = y0
    interp = [y0] * 5
    for i in range(1, len(t)):
        while t[i] > t1:
            # ---- one attempted step (rk_common._adaptive_step / _runge_kutta_step) ----
            t_new = t1 + dt
            assert t1 + dt > t1, f"underflow in dt {float(dt)}"
            assert torch.isfinite(y).all(), "non-finite values in state `y`"
            t0_s, dt_s, t1_s = t1.to(sdtype), dt.to(sdtype), t_new.to(sdtype)
            k = torch.empty(*f0.shape, 7, dtype=sdtype)
            k[..., 0] = f0
            yi = y
            for s, (alpha_s, beta_s) in enumerate(zip(alpha, beta)):
                if alpha_s == 1.0:
                    ti, perturb = t1_s, -1
                else:
                    ti, perturb = t0_s + alpha_s * dt_s, 0
                yi = y + k[..., : s + 1].matmul(beta_s * dt_s).view_as(f0)
                k[..., s + 1] = func(ti, yi, perturb=perturb)
            y_new, f_new = yi, k[..., -1]  # FSAL: c_sol == beta[-1] || 0
            y_err = k.matmul(dt_s * c_error)
            error_tol = atol + rtol * torch.max(y.abs(), y_new.abs())
            error_ratio = norm(y_err / error_tol).abs()
            stats["attempts"] += 1
            if error_ratio <= 1:
                y_mid = y + k.matmul(dt_s * c_mid).view_as(y)
                interp = _interp_fit(y, y_new, y_mid, k[..., 0], k[..., -1], dt_s)
                t0, t1, y, f0 = t1, t_new, y_new, f_new
                stats["accepted"] += 1
            dt = _optimal_step_size(dt, error_ratio)
        solution[i] = _interp_evaluate(interp, t0, t1, t[i])
    return solution


# ----------------------------------------------------------------------------------------------
# public entry points (torchdiffeq odeint / odeint_adjoint)
# ----------------------------------------------------------------------------------------------
def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, stats=None):
    """Returns the solution at `t` with the leading time axis, tuple-shaped if `y0` was a tuple."""
    if method is None:
        method = "dopri5"
    shapes = None
    if isinstance(y0, (tuple, list)):
        shapes = [y.shape for y in y0]
        rtol = _tuple_tol(rtol, shapes)
        atol = _tuple_tol(atol, shapes)
        y0 = torch.cat([y.reshape(-1) for y in y0])
        _shapes = shapes

        def norm(tensor):  # _mixed_norm over the original components
            return max(_rms_norm(c) for c in _flat_to_shape(tensor, (), _shapes))
    else:
        norm = _rms_norm
    assert t.dim() == 1 and torch.is_floating_point(t)
    reverse = bool(len(t) > 1 and (t[1:] < t[:-1]).all())
    if reverse:
        t = -t
    assert (t[1:] > t[:-1]).all(), "t must be strictly increasing or decreasing"
    f = _Func(func, shapes, reverse)
    if stats is None:
        stats = {}
    stats.update(attempts=0, accepted=0)
    with torch.no_grad():
        if method in _FIXED:
            sol = _integrate_fixed(_FIXED[method], f, y0, t)
        elif method == "dopri5":
            sol = _integrate_dopri5(f, y0, t, rtol, atol, norm, stats)
        else:
            raise ValueError(f"oracle restates euler/midpoint/rk4/dopri5 only, got {method!r}")
    stats["nfe"] = f.nfe
    if shapes is not None:
        sol = _flat_to_shape(sol, (len(t),), shapes)
    return sol


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None,
                   adjoint_params=None, **unused):
    if adjoint_params is not None:
        tuple(adjoint_params)  # the reference passes a generator; torchdiffeq materialises it
    return odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
