"""Drop-in for mdqm9/thermo/ambient/models/cpainn.py: same constructor, same `state_dict`, same
`forward(batch)` contract; the network itself runs in libtib.so."""
from __future__ import annotations

from torch import nn

from ... import _modules as M
from ..._cpainn_base import CPaiNNBase
from ...engine import Hyper


class cPaiNN(CPaiNNBase):
    """SE(3)-equivariant ChiroPaiNN drift b(t, x, T0, T1) (reference cpainn.py:10-115).

    Module positions inside `net` follow cpainn.py:67-90 so checkpoints load unchanged:
    0 spatial, 1 equivariant(zero) features, 2 edge-type embedding, 3 atom embedding,
    4/5 temperature embeddings T0/T1, 6 time embedding, 7 combine MLP, 8 PaiNNBase."""

    def __init__(self, n_features: int = 32, embedding_layers: int = 2, score_layers: int = 5, n_types=25,
                 temp_length=10, time_length=10, temperatures=[300, 400, 500, 600, 700, 800, 900, 1000]):
        super().__init__()
        self.hyper = Hyper(n_features=n_features, score_layers=score_layers, temp_length=temp_length,
                           time_length=time_length, n_types=n_types, temperatures=tuple(temperatures),
                           variant="ambient")
        self.net = nn.Sequential(
            M.Slot(),
            M.Tracked(),
            M.nominal_embedding(4, n_features),
            M.nominal_embedding(n_types, n_features),
            M.temperature_embedding(),
            M.temperature_embedding(),
            M.positional_embedding(),
            M.combine_holder(4 * n_features, n_features),
            M.painn_base(n_features, score_layers),
        )
