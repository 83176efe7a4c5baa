"""Drop-in for mdqm9/thermo/ambient/integrators.py."""
from __future__ import annotations

from typing import Optional

import torch

from .models.ode_wrapper import ODEWrapper

_FIXED_NFE = {"euler": 1, "midpoint": 2, "rk4": 4}


class MoleculeIntegrator:
    """Same constructor and `rollout(batch)` contract as the reference (integrators.py:12-68):
    returns `(xts [T,N,3], dlogp*1e2, nfe, batch.batch)` with dlogp = zeros [B], or with `return_dlogp=True`
    the integrated [T,B] change of log density (state (x, dlogp), right-hand side (b, -div*1e-2), time grid
    reversed under `reverse_ode`; integrators.py:36-53).

    `method`: 'dopri5' (adaptive, torchdiffeq semantics) or the fixed-grid 'euler' / 'midpoint' /
    'rk4' on `linspace(start, end, n_step)`.  The whole rollout - every drift evaluation and every
    state update - runs inside libtib.so on the current CUDA stream, also with `return_dlogp=True` (the tuple state
    (x, dlogp), every right-hand side the drift plus its exact divergence by forward-mode tangents).

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
        """State (x, dlogp) flattened to y = [x.reshape(-1) | dlogp]  (integrators.py:36-53), integrated inside libtib.so
        (tib_rollout_fixed_dlogp / tib_rollout_dopri5_dlogp: torchdiffeq's flattened-tuple semantics, max-of-RMS error norm,
        decreasing grids for `reverse_ode`); every right-hand side is the drift plus its exact divergence."""
        eng, pb = self.ode_wrapper.prepared(batch)
        x0 = batch.x0.to(eng.device, torch.float32)
        a, b_ = (self.end, self.start) if self.reverse_ode else (self.start, self.end)
        times = torch.linspace(a, b_, self.n_step)                          # integrators.py:41-43
        scale = self.ode_wrapper.variant_scale
        mult_b, mult_d = (-1.0, scale) if self.reverse_ode else (1.0, -scale)   # ode_wrapper.py:47-49
        if self.method not in _FIXED_NFE and self.method != "dopri5":
            raise ValueError(f"unsupported method {self.method!r}: use 'dopri5', 'euler', 'midpoint' or 'rk4'")
        xts, dlogp, stats = eng.rollout_dlogp(pb, x0, times, self.method, mult_b=mult_b, mult_d=mult_d, rtol=self.rtol,
                                              atol=self.atol, save_frames=True, norm_allreduce=self.norm_allreduce)
        nfe = stats["nfe"] if "nfe" in stats else (self.n_step - 1) * _FIXED_NFE[self.method]
        self.last_stats = dict(stats, nfe=nfe)
        if not self.save_frames:
            xts = xts[-1]
        return xts, dlogp, nfe, pb

    def rollout(self, batch, noise: Optional[torch.Tensor] = None) -> tuple:
        xts, dlogp, nfe, pb = self._solve(batch, noise)
        # one synchronisation per rollout: a tensor-core pipeline time-out or a state beyond the split-f16 range raises
        # here instead of handing garbage frames to the caller (the words are cleared once reported)
        self.ode_wrapper.b.engine().status()
        return xts, dlogp * self.dlogp_out_scale, nfe, batch.batch
