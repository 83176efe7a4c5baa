"""Drop-in for mdqm9/thermo/ambient/integrators.py."""
from __future__ import annotations

from typing import Optional

import torch

from .._tuple_ode import solve_dopri5, solve_fixed
from .models.ode_wrapper import ODEWrapper

_FIXED_NFE = {"euler": 1, "midpoint": 2, "rk4": 4}


class MoleculeIntegrator:
    """Same constructor and `rollout(batch)` contract as the reference (integrators.py:12-68):
    returns `(xts [T,N,3], dlogp*1e2, nfe, batch.batch)` with dlogp = zeros [B], or with `return_dlogp=True`
    the integrated [T,B] change of log density (state (x, dlogp), right-hand side (b, -div*1e-2), time grid
    reversed under `reverse_ode`; integrators.py:36-53).

    `method`: 'dopri5' (adaptive, torchdiffeq semantics) or the fixed-grid 'euler' / 'midpoint' /
    'rk4' on `linspace(start, end, n_step)`.  The whole rollout - every drift evaluation and every
    state update - runs inside libtib.so on the current CUDA stream.  With `return_dlogp=True` every
    right-hand side is one `tib_drift_div` call (the drift plus 3*max_atoms tangent directions) and the
    tuple-state stepper is `_tuple_ode.py` (torchdiffeq's flattened-tuple semantics, max-of-RMS error norm).

    Keyword-only extensions (not in the reference): `save_frames=False` keeps only the final state;
    `eps`, `noise`, `score` switch Euler to Euler-Maruyama with pre-drawn noise (BASELINE north_star;
    no reference oracle - with eps=0 the result is bit-identical to Euler); `norm_allreduce` lets
    shards of one batch share dopri5's global error norm; `cuda_graph=True` captures a fixed-grid rollout once per
    batch shape and replays it (the returned frames then live in a buffer that the next rollout of that shape overwrites)."""

    ode_wrapper_cls = ODEWrapper
    dlogp_out_scale = 1e2

    def __init__(self, b: torch.nn.Module, method: str = 'dopri5', n_step: int = 100, atol: float = 1e-4,
                 rtol: float = 1e-4, start: float = 0.0, end: float = 1.0, return_dlogp: bool = False,
                 reverse_ode: bool = False, *, save_frames: bool = True, eps: float = 0.0,
                 score: Optional[torch.nn.Module] = None, norm_allreduce=None, cuda_graph: bool = False) -> None:
        self.ode_wrapper = self.ode_wrapper_cls(b=b, return_dlogp=return_dlogp, reverse_ode=reverse_ode)
        self.start, self.end = start, end
        self.rtol, self.atol = rtol, atol
        self.n_step = n_step
        self.method = method
        self.return_dlogp = return_dlogp
        self.reverse_ode = reverse_ode
        self.save_frames = save_frames
        self.eps = eps
        self.score = score
        self.norm_allreduce = norm_allreduce
        self.cuda_graph = cuda_graph          # fixed-grid methods: replay the captured rollout (small batches are launch bound)
        self.last_stats = None

    def _solve(self, batch, noise=None):
        if self.return_dlogp:
            if self.eps != 0.0 or noise is not None:
                raise ValueError("Euler-Maruyama terms and return_dlogp=True are mutually exclusive")
            return self._solve_dlogp(batch)
        eng, pb = self.ode_wrapper.prepared(batch)
        x0 = batch.x0.to(eng.device, torch.float32)
        times = torch.linspace(self.start, self.end, self.n_step)        # integrators.py:56
        if self.method in _FIXED_NFE:
            score_engine = self.score.engine() if (self.score is not None and self.eps != 0.0) else None
            xts = eng.rollout_fixed(pb, x0, times, method=self.method, save_frames=self.save_frames,
                                    eps=self.eps, noise=noise, score_engine=score_engine, graph=self.cuda_graph)
            nfe = (self.n_step - 1) * _FIXED_NFE[self.method]
            self.last_stats = dict(nfe=nfe)
        elif self.method == "dopri5":
            if self.eps != 0.0 or noise is not None:
                raise ValueError("Euler-Maruyama terms need method='euler'")
            if self.end < self.start:
                raise NotImplementedError("dopri5 on a decreasing time grid is not built in the B200 path")
            xts, st = eng.rollout_dopri5(pb, x0, times, rtol=self.rtol, atol=self.atol,
                                         save_frames=self.save_frames, norm_allreduce=self.norm_allreduce)
            nfe = st["nfe"]
            self.last_stats = st
        else:
            raise ValueError(f"unsupported method {self.method!r}: use 'dopri5', 'euler', 'midpoint' or 'rk4'")
        dlogp = torch.zeros(pb.n_mol, device=eng.device)                  # integrators.py:32
        return xts, dlogp, nfe, pb

    def _solve_dlogp(self, batch):
        """State (x, dlogp) flattened to y = [x.reshape(-1) | dlogp]  (integrators.py:36-53)."""
        eng, pb = self.ode_wrapper.prepared(batch)
        x0 = batch.x0.to(eng.device, torch.float32)
        n3, n_mol = x0.numel(), pb.n_mol
        y0 = torch.cat([x0.reshape(-1), torch.zeros(n_mol, dtype=torch.float32, device=eng.device)])
        a, b_ = (self.end, self.start) if self.reverse_ode else (self.start, self.end)
        times = torch.linspace(a, b_, self.n_step)                          # integrators.py:41-43
        count = [0]

        def rhs(t, y):
            count[0] += 1
            db, dl = self.ode_wrapper.forward(t, (y[:n3].reshape(-1, 3), y[n3:]), batch)
            return torch.cat([db.reshape(-1), dl])

        stats = {}
        if self.method in _FIXED_NFE:
            sol = solve_fixed(rhs, y0, times, self.method)
        elif self.method == "dopri5":
            sol = solve_dopri5(rhs, y0, times, self.rtol, self.atol, split=n3, stats=stats)
        else:
            raise ValueError(f"unsupported method {self.method!r}: use 'dopri5', 'euler', 'midpoint' or 'rk4'")
        self.last_stats = dict(stats, nfe=count[0])
        xts = sol[:, :n3].reshape(self.n_step, -1, 3)
        dlogp = sol[:, n3:]
        if not self.save_frames:
            xts = xts[-1]
        return xts, dlogp, count[0], pb

    def rollout(self, batch, noise: Optional[torch.Tensor] = None) -> tuple:
        xts, dlogp, nfe, pb = self._solve(batch, noise)
        # one synchronisation per rollout: a tensor-core pipeline time-out or a state beyond the split-f16 range raises
        # here instead of handing garbage frames to the caller (the words are cleared once reported)
        self.ode_wrapper.b.engine().status()
        return xts, dlogp * self.dlogp_out_scale, nfe, batch.batch
