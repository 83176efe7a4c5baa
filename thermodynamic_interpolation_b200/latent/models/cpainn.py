"""Drop-in for mdqm9/thermo/latent/models/cpainn.py (noise -> data flow, one temperature `T`)."""
from __future__ import annotations

from torch import nn

from ... import _modules as M
from ..._cpainn_base import CPaiNNBase
from ...engine import Hyper


class cPaiNN(CPaiNNBase):
    """Latent ChiroPaiNN drift b(t, x, T) (reference latent cpainn.py:10-108): one temperature
    encoder when several temperatures are known (cpainn.py:43-58), none for a single-temperature
    model (cpainn.py:59-72); node ids come from `batch.atom_number`."""

    def __init__(self, n_features: int = 32, score_layers: int = 5, n_types=25, time_length=10, temp_length=10,
                 temperatures=[300, 400, 500, 600, 700, 800, 900, 1000]):
        super().__init__()
        self.hyper = Hyper(n_features=n_features, score_layers=score_layers, temp_length=temp_length,
                           time_length=time_length, n_types=n_types, temperatures=tuple(temperatures),
                           variant="latent")
        layers = [M.Slot(), M.Slot(), M.nominal_embedding(4, n_features), M.nominal_embedding(n_types, n_features)]
        if len(temperatures) > 1:
            layers += [M.temperature_embedding(), M.positional_embedding(),
                       M.combine_holder(3 * n_features, n_features)]
        else:
            layers += [M.positional_embedding(), M.combine_holder(2 * n_features, n_features)]
        layers.append(M.painn_base(n_features, score_layers))
        self.net = nn.Sequential(*layers)
