"""Reweighting statistics consumed by the reference's analysis
(mdqm9/analysis/utils/ess.py:8-10,32-35; free_energy.py:41-46): per-rank partial sums on the device
(libtib.so `tib_reweight_stats`), combined across ranks by one fp64 all-reduce (dist.py)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib

N_STATS = 5   # sum w, sum w^2, sum exp(-phi)*g, sum g, n


def reweight_partials(E0: torch.Tensor, E1: torch.Tensor, neg_dlogp: Optional[torch.Tensor] = None,
                      weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device fp64 [5]: (sum w, sum w^2, sum exp(-phi) g, sum g, n) with phi = E1 - E0 + neg_dlogp,
    w = exp(-phi) (calc_ti_weights, ess.py:8-10)."""
    dev = E0.device
    if dev.type != "cuda":
        raise RuntimeError("reweight_partials runs on CUDA tensors only (no CPU fallback)")
    f = lambda t: None if t is None else t.to(dev, torch.float64).contiguous()  # noqa: E731
    E0, E1, nd, g = f(E0), f(E1), f(neg_dlogp), f(weight)
    out = torch.empty(N_STATS, dtype=torch.float64, device=dev)
    p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tib_reweight_stats(p(E0), p(E1), p(nd), p(g), E0.numel(), out.data_ptr(),
                                                  C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "tib_reweight_stats")
    return out


def finalize(partials) -> dict:
    """ESS = (sum w)^2 / sum w^2 (calc_ESS, ess.py:32-35); dF = -log(sum exp(-phi) g / sum g)
    (calc_tfep_dF, free_energy.py:41-46)."""
    s = [float(v) for v in partials]
    n = int(round(s[4]))
    # empty input, or every weight underflowed: the statistics are undefined, not an arithmetic error
    ess = s[0] * s[0] / s[1] if s[1] > 0.0 else float("nan")
    dF = -math.log(s[2] / s[3]) if (s[2] > 0.0 and s[3] > 0.0) else float("nan")
    return dict(ess=ess, dF=dF, n=n, sum_w=s[0], sum_w2=s[1])
