"""Drop-in for adw/thermo/models/simple.py: FCNetMultiBeta, the fp64 MLP drift of the 1-D
asymmetric double well.  Same constructor and `state_dict` (`net.{0,2,..}`, `beta_embed.{0,2,4}`);
`forward` and the exact divergence run in libtib.so (csrc/adw.cuh)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from ... import _lib


class FCNetMultiBeta(nn.Module):
    def __init__(self, in_size, out_size, hidden_size, num_layers):
        super().__init__()
        if in_size != 1 or out_size != 1:
            raise ValueError("the B200 ADW kernel is built for the reference's 1-D state (in_size = out_size = 1)")
        self.hidden_size, self.num_layers = hidden_size, num_layers
        sizes = [in_size + 2] + [hidden_size] * num_layers + [out_size]    # simple.py:21
        layers = []
        for i in range(len(sizes) - 1):
            layers.append(nn.Linear(sizes[i], sizes[i + 1]))
            if i != len(sizes) - 2:
                layers.append(nn.SiLU())
        self.net = nn.Sequential(*layers)
        self.beta_embed = nn.Sequential(nn.Linear(3, hidden_size), nn.SiLU(), nn.Linear(hidden_size, hidden_size),
                                        nn.SiLU(), nn.Linear(hidden_size, 1))
        self._handle = None
        self._sig = None

    # -- engine ----------------------------------------------------------------------------------
    def _pack(self) -> np.ndarray:
        parts = []
        for seq in (self.beta_embed, self.net):
            for m in seq:
                if isinstance(m, nn.Linear):
                    parts.append(m.weight.detach().to("cpu", torch.float64).reshape(-1))
                    parts.append(m.bias.detach().to("cpu", torch.float64).reshape(-1))
        return np.ascontiguousarray(torch.cat(parts).numpy())

    def _engine(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("thermodynamic_interpolation_b200 runs on CUDA devices only (no CPU fallback)")
        sig = (str(dev),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._handle is None or sig != self._sig:
            lib = _lib.load()
            self._release()
            packed = self._pack()
            h = C.c_void_p()
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            _lib.check(lib.tib_adw_create(C.byref(h), self.hidden_size, self.num_layers,
                                          packed.ctypes.data_as(C.c_void_p), packed.size, idx), "tib_adw_create")
            self._handle, self._sig = h, sig
        return self._handle, dev

    def _release(self):
        if self._handle is not None and self._handle.value:
            _lib.load().tib_adw_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def drift_div(self, xs: torch.Tensor, t: float, beta0s: torch.Tensor, beta1s: torch.Tensor, want_div=True):
        """b(x,t,beta0,beta1) [B,1] fp64 and d b/d x [B] fp64 (unscaled)."""
        h, dev = self._engine()
        x = xs.to(dev, torch.float64).reshape(-1).contiguous()
        b0 = beta0s.to(dev, torch.float64).reshape(-1).contiguous()
        b1 = beta1s.to(dev, torch.float64).reshape(-1).contiguous()
        n = x.numel()
        out_b = torch.empty(n, dtype=torch.float64, device=dev)
        out_d = torch.empty(n, dtype=torch.float64, device=dev) if want_div else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().tib_adw_drift_div(
                h, x.data_ptr(), b0.data_ptr(), b1.data_ptr(), float(t), out_b.data_ptr(),
                out_d.data_ptr() if want_div else None, n,
                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "tib_adw_drift_div")
        return out_b.reshape(-1, 1), out_d

    def forward(self, x0s, xts, ts, beta0s, beta1s):
        """simple.py:38-41.  `ts` must be constant over the batch (it always is: ode_wrapper.py:42)."""
        t = float(ts.reshape(-1)[0])
        return self.drift_div(xts, t, beta0s, beta1s, want_div=False)[0]
