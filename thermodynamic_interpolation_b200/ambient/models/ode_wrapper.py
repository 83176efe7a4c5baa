"""Drop-in for mdqm9/thermo/ambient/models/ode_wrapper.py: the ODE right-hand side f(t, state)."""
from __future__ import annotations

import torch
from torch import nn


class ODEWrapper(nn.Module):
    """`forward(t, states, batch[, n_steps])` -> drift b(t, x) [N,3] (reference ode_wrapper.py:26-57).

    Unlike the reference it neither clones the batch nor re-stamps `batch.x` / `batch.t` per call
    (reset_batch, ode_wrapper.py:93-113): the prepared batch is cached and `x`, `t` go straight to
    the kernel.  The exact-divergence branch (`return_dlogp=True`, ode_wrapper.py:39-49,59-91) is the
    next scope row and raises here rather than silently returning zeros."""

    variant_scale = 1e-2      # ambient multiplies the divergence by 1e-2 (ode_wrapper.py:91)

    def __init__(self, b: nn.Module, return_dlogp=False, reverse_ode=False) -> None:
        super().__init__()
        self.b = b
        self.return_dlogp = return_dlogp
        self.reverse_ode = reverse_ode
        self._prepared = None      # (id(batch), PreparedBatch)

    def prepared(self, batch):
        eng = self.b.engine()
        if self._prepared is None or self._prepared[0] is not batch or self._prepared[2] is not eng:
            self._prepared = (batch, eng.prepare(batch), eng)
        return eng, self._prepared[1]

    def forward(self, integration_time, states, batch, n_steps: list = None):
        if self.return_dlogp:
            raise NotImplementedError(
                "return_dlogp=True (exact divergence, ode_wrapper.py:59-91) is not built yet in the "
                "B200 path; use return_dlogp=False")
        x = states
        if n_steps is not None:
            n_steps.append(n_steps[-1] + 1)
        eng, pb = self.prepared(batch)
        t = float(integration_time) if not torch.is_tensor(integration_time) else float(integration_time.to(torch.float32))
        return eng.drift(pb, x.to(eng.device, torch.float32), t)
