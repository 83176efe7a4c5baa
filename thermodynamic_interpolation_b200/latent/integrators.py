"""Drop-in for mdqm9/thermo/latent/integrators.py."""
from __future__ import annotations

from ..ambient.integrators import MoleculeIntegrator as _AmbientIntegrator
from .models.ode_wrapper import ODEWrapper


class MoleculeIntegrator(_AmbientIntegrator):
    """`rollout(batch) -> (xts [T,N,3], dlogp [B], batch.batch)` (latent integrators.py:41-89): a
    3-tuple, and no x1e2 rescaling of dlogp."""

    ode_wrapper_cls = ODEWrapper
    dlogp_out_scale = 1.0

    def rollout(self, batch, noise=None) -> tuple:
        xts, dlogp, _nfe, _pb = self._solve(batch, noise)
        self.ode_wrapper.b.engine().status()       # see the ambient integrator
        return xts, dlogp, batch.batch
