"""Drop-in for adw/thermo/integrators.py: StandardIntegrator for the 1-D asymmetric double well.

`rollout(x0s, beta0s, beta1s) -> (x [T,B,1], dlogp*1e2 [T,B,1])` with `return_dlogp=True`, the only
configuration that works in the reference (adw/thermo/integrators.py:33-68; its `return_dlogp=False`
branch raises a TypeError).  The right-hand side - the fp64 FCNetMultiBeta drift and
its exact 1-D divergence (adw/thermo/models/ode_wrapper.py:38-67) - is one CUDA kernel (csrc/adw.cuh);
the stepper below reproduces torchdiffeq 0.2.5 on the flattened tuple state (x, dlogp):
fixed-grid euler / midpoint / rk4 (3/8 rule) let the state drift to fp64 after the first step (fp32
state + fp64 derivative), dopri5 keeps the fp32 state and stores the fp64 derivative in fp32 stages.
The state is tiny ([2B]); the stage arithmetic runs as device tensor ops on the current stream.
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


class StandardIntegrator:
    """Same constructor as the reference (adw/thermo/integrators.py:12-31)."""

    def __init__(self, b: torch.nn.Module, method: str = 'dopri5', n_step: int = 100, atol: float = 1e-4,
                 rtol: float = 1e-4, start: float = 0.0, end: float = 1.0, return_dlogp: bool = False) -> None:
        self.b = b
        self.method, self.n_step = method, n_step
        self.atol, self.rtol = atol, rtol
        self.start, self.end = start, end
        self.return_dlogp = return_dlogp
        self.last_stats = None

    # f(t, y) on the flattened state y = [x (B) | dlogp (B)]: (b, -div * 1e-2)   (ode_wrapper.py:38-47)
    def _rhs(self, t: float, y: torch.Tensor, beta0s, beta1s) -> torch.Tensor:
        B = y.numel() // 2
        b, div = self.b.drift_div(y[:B].reshape(-1, 1), float(t), beta0s, beta1s, want_div=True)
        self._nfe += 1
        return torch.cat([b.reshape(-1), -(div * 1e-2)])

    def rollout(self, x0s: torch.Tensor, beta0s: torch.Tensor, beta1s: torch.Tensor) -> tuple:
        if not self.return_dlogp:
            raise TypeError("StandardIntegrator.rollout needs return_dlogp=True: the reference's return_dlogp=False "
                            "branch passes a 1-tuple state its wrapper cannot unpack (adw/thermo/integrators.py:57)")
        dev = next(self.b.parameters()).device
        x0 = x0s.to(dev)
        B = x0.shape[0]
        y0 = torch.cat([x0.reshape(-1), torch.zeros(B, dtype=x0.dtype, device=dev)])     # integrators.py:42
        times = torch.linspace(self.start, self.end, self.n_step)                           # fp32 grid, host
        self._nfe = 0
        self.last_stats = None
        f = lambda t, y: self._rhs(t, y, beta0s, beta1s)  # noqa: E731
        if self.method in ("euler", "midpoint", "rk4"):
            sol = self._fixed(f, y0, times)
        elif self.method == "dopri5":
            sol = self._dopri5(f, y0, times)
        else:
            raise ValueError(f"unsupported method {self.method!r}")
        self.last_stats = dict(self.last_stats or {}, nfe=self._nfe)
        x = sol[:, :B].reshape(self.n_step, B, 1)
        dlogp = sol[:, B:].reshape(self.n_step, B, 1)
        return x, dlogp * 1e2                                                                # integrators.py:68

    # torchdiffeq FixedGridODESolver.integrate: y1 = y0 + step; the output grid is the step grid
    def _fixed(self, f, y0, times):
        sol = torch.empty((len(times),) + tuple(y0.shape), dtype=y0.dtype, device=y0.device)
        sol[0] = y0
        y = y0
        third, two_thirds = 1 / 3, 2 / 3
        for j in range(1, len(times)):
            t0, t1 = times[j - 1], times[j]
            dt = t1 - t0                                   # fp32 tensor arithmetic, as in torchdiffeq
            dtf = float(dt)
            if self.method == "euler":
                inc = dtf * f(float(t0), y)
            elif self.method == "midpoint":
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
    def _dopri5(self, f, y0, times):
        B = y0.numel() // 2
        sd = y0.dtype
        dev = y0.device
        t = times.to(torch.float64)
        rtol, atol = float(self.rtol), float(self.atol)

        def norm(v):                                       # max of the per-component RMS norms (x, dlogp)
            v = v.to(torch.float64)
            return float(torch.maximum(v[:B].pow(2).mean().sqrt(), v[B:].pow(2).mean().sqrt()))

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
        self.last_stats = dict(attempts=attempts, accepted=accepted)
        return sol
