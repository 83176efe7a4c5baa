"""Drop-in for adw/thermo/integrators.py: StandardIntegrator for the 1-D asymmetric double well.

`rollout(x0s, beta0s, beta1s) -> (x [T,B,1], dlogp*1e2 [T,B,1])` with `return_dlogp=True`, the only
configuration that works in the reference (adw/thermo/integrators.py:33-68; its `return_dlogp=False`
branch raises a TypeError).  The right-hand side - the fp64 FCNetMultiBeta drift and
its exact 1-D divergence (adw/thermo/models/ode_wrapper.py:38-67) - is one CUDA kernel (csrc/adw.cuh);
the steppers in `_tuple_ode.py` reproduce torchdiffeq 0.2.5 on the flattened tuple state (x, dlogp):
fixed-grid euler / midpoint / rk4 (3/8 rule) let the state drift to fp64 after the first step (fp32
state + fp64 derivative), dopri5 keeps the fp32 state and stores the fp64 derivative in fp32 stages.
The state is tiny ([2B]); the stage arithmetic runs as device tensor ops on the current stream.
"""
from __future__ import annotations

import torch

from .._tuple_ode import solve_dopri5, solve_fixed


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
            sol = solve_fixed(f, y0, times, self.method)
        elif self.method == "dopri5":
            st = {}
            sol = solve_dopri5(f, y0, times, self.rtol, self.atol, split=B, stats=st)
            self.last_stats = st
        else:
            raise ValueError(f"unsupported method {self.method!r}")
        self.last_stats = dict(self.last_stats or {}, nfe=self._nfe)
        x = sol[:, :B].reshape(self.n_step, B, 1)
        dlogp = sol[:, B:].reshape(self.n_step, B, 1)
        return x, dlogp * 1e2                                                                # integrators.py:68
