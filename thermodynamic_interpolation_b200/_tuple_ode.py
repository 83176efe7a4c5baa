"""Host-side ODE steppers for tuple states (x, dlogp), shared by the ADW StandardIntegrator and by the
molecule integrators when `return_dlogp=True`.

They restate what torchdiffeq 0.2.5 does with the flattened tuple state y = [x.reshape(-1) | dlogp]
(reference call sites: adw/thermo/integrators.py:49-55, mdqm9/thermo/ambient/integrators.py:45-53,
latent/integrators.py:66-74): fixed-grid euler / midpoint / rk4 (3/8 rule) on the output grid with the
time arithmetic in the grid dtype, and dopri5 with time in fp64, state and stages in the state dtype,
stage times cast to the state dtype, alpha == 1 stages at nextafter(t1, -inf), the error ratio as the
MAX of the per-component RMS norms, the 0.9 / 0.2 / 10 controller, Hairer's initial step, no clipping
to the end time and quartic dense output.  The right-hand side (one CUDA call per evaluation - the
drift with its exact divergence) dominates by orders of magnitude; the state arithmetic here is a
handful of small device tensor ops on the current stream.
"""
from __future__ import annotations

import torch

_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_C_ERR = [c - d for c, d in zip(_C_SOL, [1951 / 21600, 0, 22642 / 50085, 451 / 720, -12231 / 42400, 649 / 6300, 1 / 60])]
_C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
          187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]



# torchdiffeq FixedGridODESolver.integrate: y1 = y0 + step; the output grid is the step grid
def solve_fixed(f, y0, times, method):
    sol = torch.empty((len(times),) + tuple(y0.shape), dtype=y0.dtype, device=y0.device)
    sol[0] = y0
    y = y0
    third, two_thirds = 1 / 3, 2 / 3
    for j in range(1, len(times)):
        t0, t1 = times[j - 1], times[j]
        dt = t1 - t0                                   # fp32 tensor arithmetic, as in torchdiffeq
        dtf = float(dt)
        if method == "euler":
            inc = dtf * f(float(t0), y)
        elif method == "midpoint":
            half = float(0.5 * dt)
            y_mid = y + f(float(t0), y) * half
            inc = dtf * f(float(t0 + 0.5 * dt), y_mid)
        else:                                          # rk4 = 3/8 rule (rk_common.rk4_alt_step_func)
            k1 = f(float(t0), y)
            k2 = f(float(t0 + dt * third), y + dtf * k1 * third)
            k3 = f(float(t0 + dt * two_thirds), y + dtf * (k2 - k1 * third))
            k4 = f(float(t1), y + dtf * (k1 - k2 + k3))
            inc = (k1 + 3 * (k2 + k3) + k4) * dtf * 0.125
        y = y + inc                                    # fp32 + fp64 -> fp64 from the first step on
        sol[j] = y
    return sol


# torchdiffeq RKAdaptiveStepsizeODESolver (dopri5): fp32 state and stages, fp64 time, mixed RMS norm
def solve_dopri5(f, y0, times, rtol, atol, split, stats=None):
    """`split` = number of leading elements of y that form the first tuple component."""
    sd = y0.dtype
    dev = y0.device
    t = times.to(torch.float64)
    rtol, atol = float(rtol), float(atol)
    if len(t) > 1 and float(t[-1]) < float(t[0]):       # decreasing grid: solve (-t, -f) (torchdiffeq _check_inputs)
        g = f
        f = lambda tt, y: -g(-tt, y)  # noqa: E731
        t = -t

    def norm(v):                                       # max of the per-component RMS norms (x, dlogp)
        v = v.to(torch.float64)
        return float(torch.maximum(v[:split].pow(2).mean().sqrt(), v[split:].pow(2).mean().sqrt()))

    def fs(tt, y, perturb=0):
        ts = torch.tensor(tt, dtype=torch.float64).to(sd)         # _PerturbFunc casts t to the state dtype
        if perturb < 0:
            ts = torch.nextafter(ts, ts - 1)
        return f(float(ts), y)

    beta = [torch.tensor(b, dtype=torch.float64).to(sd).to(dev) for b in _BETA]
    c_err = torch.tensor(_C_ERR, dtype=torch.float64).to(sd).to(dev)
    c_mid = torch.tensor(_C_MID, dtype=torch.float64).to(sd).to(dev)
    sol = torch.empty((len(t),) + tuple(y0.shape), dtype=sd, device=dev)
    sol[0] = y0
    f0 = fs(float(t[0]), y0)
    # misc._select_initial_step, order 4
    scale = atol + y0.abs() * rtol
    d0, d1 = norm(y0 / scale), norm(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = fs(float(t[0]) + h0, y0 + h0 * f0)
    d2 = abs(norm((f1 - f0) / scale) / h0)
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    dt = min(100 * h0, h1)
    t0 = t1 = float(t[0])
    y = y0
    interp = None
    attempts = accepted = 0
    for i in range(1, len(t)):
        ti_out = float(t[i])
        while ti_out > t1:
            t_new = t1 + dt
            if not t1 + dt > t1:
                raise RuntimeError(f"underflow in dt {dt}")
            t0_s = torch.tensor(t1, dtype=torch.float64).to(sd)
            dt_s = torch.tensor(dt, dtype=torch.float64).to(sd)
            t1_s = torch.tensor(t_new, dtype=torch.float64).to(sd)
            k = torch.empty((y.numel(), 7), dtype=sd, device=dev)
            k[:, 0] = f0
            yi = y
            for s in range(6):
                yi = y + k[:, : s + 1].matmul(beta[s] * dt_s.to(dev))
                if _ALPHA[s] == 1.0:
                    k[:, s + 1] = fs(float(t1_s), yi, perturb=-1)
                else:
                    k[:, s + 1] = fs(float(t0_s + torch.tensor(_ALPHA[s], dtype=torch.float64).to(sd) * dt_s), yi)
            y_new, f_new = yi, k[:, 6]
            y_err = k.matmul(dt_s.to(dev) * c_err)
            tol = atol + rtol * torch.maximum(y.abs(), y_new.abs()).to(torch.float64)
            ratio = norm(y_err.to(torch.float64) / tol)
            attempts += 1
            if ratio != ratio:
                raise RuntimeError("dopri5: non-finite error ratio")
            if ratio <= 1:
                y_mid = y + k.matmul(dt_s.to(dev) * c_mid)
                dts = dt_s.to(dev)
                fa, fb = k[:, 0], k[:, 6]
                interp = (y, dts * fa,
                          dts * (fb - 4 * fa) - 11 * y - 5 * y_new + 16 * y_mid,
                          dts * (5 * fa - 3 * fb) + 18 * y + 14 * y_new - 32 * y_mid,
                          2 * dts * (fb - fa) - 8 * (y_new + y) + 16 * y_mid)
                t0, t1, y, f0 = t1, t_new, y_new, f_new
                accepted += 1
            if ratio == 0:
                dt = dt * 10.0
            else:
                dfactor = 1.0 if ratio < 1 else 0.2
                dt = dt * min(10.0, max(0.9 / ratio ** 0.2, dfactor))
        xx = torch.tensor((ti_out - t0) / (t1 - t0), dtype=torch.float64).to(sd).to(dev)
        total = interp[0] + xx * interp[1]
        xp = xx
        for c in interp[2:]:
            xp = xp * xx
            total = total + xp * c
        sol[i] = total
    if stats is not None:
        stats.update(attempts=attempts, accepted=accepted)
    return sol
