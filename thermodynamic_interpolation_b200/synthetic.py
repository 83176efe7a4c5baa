"""Deterministic random-init weights for the synthetic workloads (bench.py, tools/, tests): there are no trained
checkpoints offline (the reference's are on Zenodo, README.md:32), so every measured network is the reference
architecture under `torch.manual_seed(seed)` with seeded noise on every parameter."""
from __future__ import annotations

import torch


def perturb_(module: torch.nn.Module, seed: int, scale: float = 0.05) -> torch.nn.Module:
    """Adds seeded Gaussian noise to every floating parameter (in `parameters()` order) so that
    LayerNorm gains/offsets and biases are not at their trivial defaults.  Scalar parameters (the
    reference's `device_tracker` dummies) are skipped."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            if p.dim() == 0:
                continue
            noise = torch.randn(p.shape, generator=gen, dtype=torch.float32)
            p.add_(noise.to(p.dtype) * scale)
    return module


def seeded_ambient_model(n_features: int = 128, score_layers: int = 5, temp_length: float = 100, seed: int = 0):
    """The cfg-2 / cfg-4 drift network: cPaiNN(F, L) under manual_seed(seed), perturbed with seed + 1 (the recipe of
    oracle/make_golden.py, so the reference's own class yields bit-identical weights)."""
    from .ambient.models.cpainn import cPaiNN
    torch.manual_seed(seed)
    return perturb_(cPaiNN(n_features=n_features, score_layers=score_layers, temp_length=temp_length), seed + 1).eval()
