"""Drop-in for mdqm9/thermo/ambient/losses.py: `StandardVelocityLoss(interpolant, t_distr)(batch0, batch1, b)` returns
the scalar loss of the reference (losses.py:30-85,126-133) and `loss.backward()` fills `.grad` of the model's
parameters, so the reference's loop (`optim.zero_grad(); loss = loss_fn(...); loss.backward(); clip_grad_norm_;
optim.step()`, mdqm9/train_ambient.py:124-148) runs unchanged.  Loss AND gradients are computed by one call into
libtib.so (tib_train_loss_grad: forward of both antithetic passes and hand-written adjoints on the device); autograd
only carries the finished gradient vector to the parameters.  No eager / autograd fallback exists."""
from __future__ import annotations

import torch

from ..train import TrainEngine, flatten, packed_parameters


def draw_times(n_atoms_per_mol, t_distr: str = "uniform") -> torch.Tensor:
    """One draw per molecule from torch's CPU generator, repeated over its atoms - the reference's order of draws
    (losses.py:46-50).  Returns t [N,1]."""
    if t_distr == "uniform":
        return torch.cat([torch.rand(1).repeat(n) for n in n_atoms_per_mol]).unsqueeze(1)
    if t_distr == "beta":
        dist = torch.distributions.beta.Beta(0.5, 0.5)
        return torch.cat([dist.sample((1,)).repeat(n) for n in n_atoms_per_mol]).unsqueeze(1)
    raise ValueError(f"Invalid value of time distribution: {t_distr}")


class _NativeLoss(torch.autograd.Function):
    """Carries the library's gradient vector to the parameters."""

    @staticmethod
    def forward(ctx, run, *params):
        loss, flat_grad = run(params)
        ctx.flat_grad = flat_grad
        ctx.shapes = [p.shape for p in params]
        return loss

    @staticmethod
    def backward(ctx, gout):
        grads, off = [], 0
        for shp in ctx.shapes:
            n = shp.numel()
            grads.append((ctx.flat_grad[off:off + n] * gout).view(shp))
            off += n
        return (None, *grads)


class BaseVelocityLoss(torch.nn.Module):
    def __init__(self, interpolant, t_distr: str = "uniform") -> None:
        super().__init__()
        assert t_distr in ["uniform", "beta"], f"Invalid time distribution: {t_distr}"
        self.interpolant = interpolant
        self.t_distr = t_distr
        self._engine = None

    def engine(self, b) -> TrainEngine:
        if self._engine is None or self._engine.hp != b.hyper or self._engine.device != b.device:
            self._engine = TrainEngine(b.hyper, b.device)
        return self._engine

    def forward(self, batch0, batch1, b, t=None, z=None) -> torch.Tensor:
        """`t` [N,1] / `z` [N,3] override the draws (tests, reproducing a reference run on identical noise)."""
        kind = getattr(self.interpolant, "kind", None)
        if kind not in ("brownian", "sin2"):
            raise NotImplementedError(f"the native loss supports LinearInterpolant with gamma 'brownian' or 'sin2', got {kind!r}")
        eng = self.engine(b)
        tb = eng.prepare(batch0, batch1)
        if t is None:
            t = draw_times(tb.n_atoms, self.t_distr)
        if z is None:
            z = torch.randn(tb.x0.shape)                  # interpolants.py:29, CPU generator
        params = packed_parameters(b)

        def run(ps):
            loss, grad, _ = eng.loss_and_grad(flatten(ps), tb, t, z, gamma=kind, a=self.interpolant.a_value)
            eng.status()
            return loss.to(torch.float32).reshape(()), grad

        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _NativeLoss.apply(run, *params)
        return run(params)[0]


class StandardVelocityLoss(BaseVelocityLoss):
    """losses.py:119-133: 0.5 |b+|^2 - (dtI + gamma_dot z) . b+ + 0.5 |b-|^2 - (dtI - gamma_dot z) . b-, mean over atoms."""
