"""Drop-in for mdqm9/thermo/ambient/interpolants.py: the same classes, constructor arguments and callables
(`gamma`, `gamma_dot`, `gg_dot`, `It`, `dtIt`, `calc_antithetic_xts`, `calc_regular_xt`).  The training loss does not
call these tensor lambdas - it hands (`gamma_kind`, `a`) to libtib.so, which evaluates the interpolant on the device
(csrc/train.cuh::k_tr_interp) - they exist so that code written against the reference's interpolant API keeps working."""
from __future__ import annotations

import torch
import torch.nn as nn


class BaseInterpolant(nn.Module):
    """interpolants.py:5-50."""

    def __init__(self):
        super().__init__()
        self.gamma = None
        self.It = None

    def calc_antithetic_xts(self, t, x0, x1):
        """(x_t^+, x_t^-, z) with z ~ N(0, 1) drawn from torch's CPU generator (interpolants.py:16-33)."""
        z = torch.randn(x0.shape).to(t)
        gamma, It = self.gamma(t), self.It(t, x0, x1)
        return It + gamma * z, It - gamma * z, z

    def calc_regular_xt(self, t, x0, x1):
        z = torch.randn(x0.shape).to(t)
        return self.It(t, x0, x1) + self.gamma(t) * z, z

    def forward(self):
        raise NotImplementedError


class LinearInterpolant(BaseInterpolant):
    """I_t = (1 - t) x0 + t x1 with gamma in {'brownian', 'sin2', 'sig_sum'} (interpolants.py:53-108).  `kind` / `a_value`
    are what the native loss consumes; 'sig_sum' has no native kernel."""

    def __init__(self, a: float = 1, gamma: str = "brownian") -> None:
        super().__init__()
        self.kind, self.a_value = gamma, float(a)
        if gamma == "brownian":
            at = torch.tensor(a)
            self.gamma = lambda t: torch.sqrt(at * t * (1 - t))
            self.gamma_dot = lambda t: (1 / (2 * torch.sqrt(at * t * (1 - t)))) * at * (1 - 2 * t)
            self.gg_dot = lambda t: (at / 2) * (1 - 2 * t)
        elif gamma == "sin2":
            self.gamma = lambda t: torch.sin(torch.pi * t) ** 2
            self.gamma_dot = lambda t: 2 * torch.pi * torch.sin(torch.pi * t) * torch.cos(torch.pi * t)
            self.gg_dot = lambda t: self.gamma(t) * self.gamma_dot(t)
        elif gamma == "sig_sum":
            at, sf = torch.tensor(a), torch.tensor(2.2)
            sg = torch.sigmoid
            self.gamma = lambda t: sf * (sg(at * (t - 0.5) + 1) - sg(at * (t - 0.5) - 1) - sg(-at / 2 + 1) + sg(-at / 2 - 1))
            self.gamma_dot = lambda t: sf * (-at * (1 - sg(-1 + at * (t - 0.5))) * sg(-1 + at * (t - 0.5))
                                             + at * (1 - sg(1 + at * (t - 0.5))) * sg(1 + at * (t - 0.5)))
            self.gg_dot = lambda t: self.gamma(t) * self.gamma_dot(t)
        else:
            raise NotImplementedError
        self.a = lambda t: (1 - t)
        self.adot = lambda t: -1.0
        self.b = lambda t: t
        self.bdot = lambda t: 1.0
        self.It = lambda t, x0, x1: self.a(t) * x0 + self.b(t) * x1
        self.dtIt = lambda t, x0, x1: self.adot(t) * x0 + self.bdot(t) * x1
