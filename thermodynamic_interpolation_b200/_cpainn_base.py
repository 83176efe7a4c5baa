"""Shared forward plumbing of the ambient and latent cPaiNN drop-ins."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .engine import DriftEngine, Hyper


class CPaiNNBase(nn.Module):
    """`forward(batch) -> batch` with `batch.output [N,3]`, computed by libtib.so.

    The engine (weights repacked on the device) is built lazily on first use and rebuilt whenever a
    parameter changed (`load_state_dict`, `.to(device)`, an optimiser step)."""

    hyper: Hyper

    def __init__(self):
        super().__init__()
        self._engine: Optional[DriftEngine] = None
        self._engine_sig = None
        self._math_mode = None     # None = the library default (tensor cores when n_features is 128 or 256)

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def _signature(self):
        return (str(self.device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self) -> DriftEngine:
        sig = self._signature()
        if self._engine is None or sig != self._engine_sig:
            self._engine = DriftEngine(self.state_dict(), self.hyper, self.device)
            if self._math_mode is not None:
                self._engine.set_math(self._math_mode)
            self._engine_sig = sig
        return self._engine

    def set_math(self, mode: int):
        """Select the GEMM arithmetic (thermodynamic_interpolation_b200._lib.MATH_*)."""
        self._math_mode = int(mode)
        if self._engine is not None:
            self._engine.set_math(self._math_mode)
        return self

    def forward(self, batch):
        """cPaiNN.forward (cpainn.py:93-115): reads batch.x and batch.t, writes batch.output."""
        eng = self.engine()
        pb = eng.prepare(batch)
        t = batch.t
        t_val = float(t.reshape(-1)[0]) if torch.is_tensor(t) else float(t)
        x = batch.x.to(eng.device, torch.float32)
        batch.output = eng.drift(pb, x, t_val)
        eng.status()      # synchronises the stream and raises on a recorded tensor-core pipeline / range fault
        return batch
