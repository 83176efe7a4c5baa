"""Drop-in for mdqm9/thermo/latent/models/ode_wrapper.py."""
from __future__ import annotations

from ...ambient.models.ode_wrapper import ODEWrapper as _AmbientODEWrapper


class ODEWrapper(_AmbientODEWrapper):
    """Same RHS; the latent wrapper takes no `n_steps` list (latent ode_wrapper.py:29) and leaves the
    divergence unscaled (latent ode_wrapper.py:86)."""

    variant_scale = 1.0

    def forward(self, integration_time, states, batch):
        return super().forward(integration_time, states, batch, None)
