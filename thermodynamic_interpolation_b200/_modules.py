"""Parameter holders that reproduce the reference's module tree - and therefore its `state_dict`
key set, key order and default initialisation order - without any of its compute.

The reference checkpoints are plain `state_dict`s (mdqm9/sample_ambient.py:131); the weight interface
of the drop-in is that key set, including the scalar `device_tracker` dummies of
mdqm9/thermo/ambient/models/device.py:18-26.  All arithmetic happens in libtib.so.
"""
from __future__ import annotations

import torch
from torch import nn


class Slot(nn.Module):
    """A parameter-free position in `cPaiNN.net` (AddSpatialFeatures, latent AddEquivariantFeatures)."""


class Tracked(nn.Module):
    """Carries the reference's `device_tracker` dummy parameter (device.py:25)."""

    def __init__(self):
        super().__init__()
        self.device_tracker = nn.Parameter(torch.tensor(1.0))


def mlp_holder(f_in: int, f_hidden: int, f_out: int) -> nn.Module:
    """Keys `<name>.mlp.{0,1,3,4,6}.*` of embedding.MLP (embedding.py:26-34)."""
    holder = nn.Module()
    holder.mlp = nn.Sequential(
        nn.Linear(f_in, f_hidden), nn.LayerNorm(f_hidden), nn.SiLU(),
        nn.Linear(f_hidden, f_hidden), nn.LayerNorm(f_hidden), nn.SiLU(),
        nn.Linear(f_hidden, f_out))
    return holder


def nominal_embedding(n_types: int, n_features: int) -> nn.Module:
    """Key `<name>.embedding.weight` of NominalEmbedding (embedding.py:89-103)."""
    holder = nn.Module()
    holder.embedding = nn.Embedding(n_types, n_features)
    return holder


def positional_embedding() -> nn.Module:
    """Key `<name>.embedding.device_tracker` of PositionalEmbedding (embedding.py:163-181)."""
    holder = nn.Module()
    holder.embedding = Tracked()
    return holder


def temperature_embedding() -> nn.Module:
    """Keys `<name>.embedding.device_tracker`, `<name>.embedding.positional_encoding.device_tracker`
    of TemperatureEmbedding / TemperatureEncoder (embedding.py:184-230)."""
    holder = nn.Module()
    enc = Tracked()
    enc.positional_encoding = Tracked()
    holder.embedding = enc
    return holder


def combine_holder(f_in: int, f_out: int) -> nn.Module:
    """CombineInvariantFeatures (embedding.py:233-247): `<name>.mlp.mlp.*`."""
    holder = nn.Module()
    holder.mlp = mlp_holder(f_in, f_out, f_out)
    return holder


def equivariant_linear(f_in: int, f_out: int) -> nn.Module:
    """EquivariantLinear (cpainn.py:379-390): `<name>.linear.weight`, no bias."""
    holder = nn.Module()
    holder.linear = nn.Linear(f_in, f_out, bias=False)
    return holder


def painn_base(n_features: int, n_layers: int) -> nn.Module:
    """PaiNNBase (cpainn.py:118-150): layers.{2l} = SE3Message, layers.{2l+1} = Update,
    layers.{2L} = LayerReadout(n_features, 1)."""
    F = n_features
    layers = []
    for _ in range(n_layers):
        msg = nn.Module()                      # SE3Message, cpainn.py:254-261
        msg.positional_encoder = Tracked()
        msg.phi = mlp_holder(2 * F, F, 5 * F)
        msg.w = mlp_holder(F, F, 5 * F)
        layers.append(msg)
        upd = nn.Module()                      # Update, cpainn.py:339-343
        upd.u = equivariant_linear(F, F)
        upd.v = equivariant_linear(F, F)
        upd.mlp = mlp_holder(2 * F, F, 3 * F)
        layers.append(upd)
    ro = nn.Module()                           # LayerReadout, cpainn.py:418-423
    ro.mlp = mlp_holder(F, F, 2)
    ro.V = equivariant_linear(F, 1)
    layers.append(ro)
    base = nn.Module()
    base.layers = nn.Sequential(*layers)
    return base
