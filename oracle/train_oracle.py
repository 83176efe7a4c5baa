"""TEST INFRASTRUCTURE ONLY (oracle/).  CPU restatement of the reference's training step for the
ambient drift network (BASELINE configs[4]):

  * LinearInterpolant                       mdqm9/thermo/ambient/interpolants.py:53-108
  * BaseInterpolant.calc_antithetic_xts     mdqm9/thermo/ambient/interpolants.py:16-33
  * BaseVelocityLoss.forward                mdqm9/thermo/ambient/losses.py:30-85
  * StandardVelocityLoss.loss_per_sample    mdqm9/thermo/ambient/losses.py:126-133
  * clip_grad_norm_(params, 1) + Adam       mdqm9/train_ambient.py:96,144-148  (torch 2.6 semantics)

The drift is oracle/cpainn_oracle.py::drift evaluated with per-node times (losses.py:69-70 sets
batch.t = t.squeeze(), one uniform draw per molecule repeated over its atoms); gradients come from
torch autograd through that restatement.  Pinned against tests/golden/train_*.npz, which
oracle/make_golden.py::train_case produces with the UNMODIFIED reference loss, backward, clipping
and torch.optim.Adam (tests/test_train_oracle.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from . import cpainn_oracle as co


# ---- interpolants.py:53-108 -------------------------------------------------------------------------
def gamma_fns(kind: str, a: float = 1.0):
    """(gamma, gamma_dot) of LinearInterpolant(a, gamma=kind)."""
    if kind == "brownian":
        at = torch.tensor(a)
        return (lambda t: torch.sqrt(at * t * (1 - t)),
                lambda t: (1 / (2 * torch.sqrt(at * t * (1 - t)))) * at * (1 - 2 * t))
    if kind == "sin2":
        return (lambda t: torch.sin(torch.pi * t) ** 2,
                lambda t: 2 * torch.pi * torch.sin(torch.pi * t) * torch.cos(torch.pi * t))
    raise NotImplementedError(kind)


def draw_t_z(n_atoms_per_mol: List[int], t_distr: str = "uniform") -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's draws from the global CPU generator, in its order: one `torch.rand(1)` (or
    Beta(0.5, 0.5) sample) per molecule (losses.py:46-50), then `torch.randn(x0.shape)`
    (interpolants.py:29).  Returns t [N,1], z [N,3]."""
    if t_distr == "uniform":
        t = torch.cat([torch.rand(1).repeat(n) for n in n_atoms_per_mol]).unsqueeze(1)
    elif t_distr == "beta":
        dist = torch.distributions.beta.Beta(0.5, 0.5)
        t = torch.cat([dist.sample((1,)).repeat(n) for n in n_atoms_per_mol]).unsqueeze(1)
    else:
        raise ValueError(t_distr)
    z = torch.randn(sum(n_atoms_per_mol), 3)
    return t, z


def velocity_loss(sd: Dict[str, torch.Tensor], hp: co.Hyper, x0, x1, t, z, atoms, edge_index, edge_type, T0, T1,
                  gamma: str = "sin2", a: float = 1.0, detach: bool = False):
    """StandardVelocityLoss(LinearInterpolant(a, gamma))(batch0, batch1, b) for given draws (t [N,1], z [N,3]).
    Returns (loss scalar, b_plus, b_minus)."""
    g, gdot = gamma_fns(gamma, a)
    It = (1 - t) * x0 + t * x1                                  # interpolants.py:100-104
    xtp, xtm = It + g(t) * z, It - g(t) * z                     # interpolants.py:29-33
    xtp = xtp - torch.mean(xtp, dim=0)                          # losses.py:56-57: mean over ALL atoms of the batch
    xtm = xtm - torch.mean(xtm, dim=0)
    tn = t.squeeze(1)
    bp = co.drift(sd, hp, xtp, tn, atoms, edge_index, edge_type, T0=T0, T1=T1, detach=detach)
    bm = co.drift(sd, hp, xtm, tn, atoms, edge_index, edge_type, T0=T0, T1=T1, detach=detach)
    dtIt = -1.0 * x0 + 1.0 * x1                                 # interpolants.py:97-104
    gd = gdot(t)
    per_atom = 0.5 * (bp ** 2).sum(1) - ((dtIt + gd * z) * bp).sum(1)      # losses.py:126-133, vmapped over rows
    per_atom = per_atom + 0.5 * (bm ** 2).sum(1) - ((dtIt - gd * z) * bm).sum(1)
    return per_atom.mean(), bp, bm


def loss_and_grads(sd, hp, *args, **kw):
    """Loss and d loss / d weight for every floating-point tensor of the state_dict (device_tracker dummies get None,
    like in the reference, where they do not take part in the forward)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    loss, bp, bm = velocity_loss(leaves, hp, *args, detach=False, **kw)
    names = list(leaves)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return loss.detach(), {k: g for k, g in zip(names, grads)}, bp.detach(), bm.detach()


# ---- train_ambient.py:144-148 -------------------------------------------------------------------------
def clip_grad_norm(grads: Dict[str, torch.Tensor], max_norm: float = 1.0) -> float:
    """torch.nn.utils.clip_grad_norm_ (2-norm): scales in place by min(1, max_norm / (total + 1e-6)); returns total."""
    gs = [g for g in grads.values() if g is not None]
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in gs]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in gs:
        g.mul_(coef)
    return float(total)


class Adam:
    """torch.optim.Adam (no amsgrad, L2 weight decay added to the gradient): the single-tensor update rule."""

    def __init__(self, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.step_no = 0
        self.m: Dict[str, torch.Tensor] = {}
        self.v: Dict[str, torch.Tensor] = {}

    def step(self, params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor]):
        self.step_no += 1
        b1, b2 = self.betas
        bc1, bc2 = 1 - b1 ** self.step_no, 1 - b2 ** self.step_no
        for k, g in grads.items():
            if g is None:
                continue
            p = params[k]
            if self.wd:
                g = g + self.wd * p
            m = self.m.setdefault(k, torch.zeros_like(p))
            v = self.v.setdefault(k, torch.zeros_like(p))
            m.lerp_(g, 1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-self.lr / bc1)
