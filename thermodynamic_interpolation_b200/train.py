"""Host-side driver of the training entry points of libtib.so (tib_train_loss_grad, tib_adam_step; include/tib.h):
flat device weight vectors in the C ABI's packing order, batch preparation, workspace.  PyTorch is used for device
memory, streams and torch.distributed only - every arithmetic step of the loss, its gradients and the optimiser runs
in the library.  There is no autograd or eager fallback."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib
from .batch import MolBatch
from .engine import Hyper, PreparedBatch, model_desc, packed_keys


def packed_parameters(model) -> List[torch.nn.Parameter]:
    """The model's parameters in the packing order of tib_packed_weight_count."""
    named = dict(model.named_parameters())
    return [named[k] for k, _ in packed_keys(model.hyper)]


def flatten(params) -> torch.Tensor:
    return torch.cat([p.detach().reshape(-1).to(torch.float32) for p in params])


class PreparedTrainBatch:
    """batch0 / batch1 of the training loop (contract of MDQM9MultiTempDataset.process, mdqm9/data/mdqm9_ambient.py:87-107:
    x, T, atoms, edge_index, edge_type, batch) as device arrays in the layout `tib_train_batch` wants."""

    _ARRAYS = ("mol_ptr", "edge_ptr", "atom_id", "edge_type", "temp0", "temp1", "batch_vec")

    def __init__(self, batch0, batch1, hp: Hyper, device, validate: bool = True):
        view = MolBatch(atoms=batch0.atoms, batch=batch0.batch, edge_index=batch0.edge_index, edge_type=batch0.edge_type,
                        T0=batch0.T, T1=batch1.T)
        device = torch.device(device)
        if batch0.atoms.device.type == "cpu" and device.type == "cuda":
            # host batches (what a DataLoader yields): validate and derive the index arrays on the host - no device
            # synchronisation, so the preparation of step k + 1 overlaps the device work of step k - then copy
            self.pb = PreparedBatch(view, hp, torch.device("cpu"), validate, sampling_tables=False)
            for name in self._ARRAYS:
                setattr(self.pb, name, getattr(self.pb, name).to(device, non_blocking=True))
        else:
            self.pb = PreparedBatch(view, hp, device, validate, sampling_tables=False)
        self.x0 = batch0.x.to(device, torch.float32, non_blocking=True).contiguous()
        self.x1 = batch1.x.to(device, torch.float32, non_blocking=True).contiguous()
        if self.x0.shape != (self.pb.n_nodes, 3) or self.x1.shape != self.x0.shape:
            raise ValueError("batch0.x and batch1.x must both be [N,3] for the same molecules")
        self._n_atoms = None

    @property
    def n_atoms(self):
        """Atoms per molecule (host list; only needed to draw the per-molecule times)."""
        if self._n_atoms is None:
            self._n_atoms = torch.diff(self.pb.mol_ptr).tolist()
        return self._n_atoms


class TrainEngine:
    """Loss + gradients of the reference's training step for one hyper-parameter set on one device."""

    def __init__(self, hp: Hyper, device):
        self.lib = _lib.load()
        self.hp = hp
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("thermodynamic_interpolation_b200 trains on CUDA devices only (no CPU fallback); "
                               f"got device {self.device}")
        if hp.variant != "ambient":
            raise RuntimeError("only the ambient drift network has a training step (mdqm9/train_ambient.py)")
        self.desc = model_desc(hp)
        self.n_weights = int(self.lib.tib_packed_weight_count(C.byref(self.desc)))
        self._ws: Optional[torch.Tensor] = None
        self._scratch = torch.zeros(1, dtype=torch.float64, device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def prepare(self, batch0, batch1, validate: bool = True) -> PreparedTrainBatch:
        return PreparedTrainBatch(batch0, batch1, self.hp, self.device, validate)

    def _workspace(self, tb: PreparedTrainBatch):
        need = int(self.lib.tib_train_workspace_bytes(C.byref(self.desc), tb.pb.n_mol, tb.pb.n_nodes, tb.pb.n_edges))
        if self._ws is None or self._ws.numel() < need + 256:
            self._ws = None
            free, _ = torch.cuda.mem_get_info(self.device)
            if need + 256 > free:
                raise RuntimeError(f"the training workspace for {tb.pb.n_mol} molecules is {need / 2**30:.1f} GiB but only "
                                   f"{free / 2**30:.1f} GiB are free on {self.device}: use smaller batches")
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        p = self._ws.data_ptr()
        off = (-p) % 256
        return p + off, self._ws.numel() - off

    def loss_and_grad(self, weights: torch.Tensor, tb: PreparedTrainBatch, t: torch.Tensor, z: torch.Tensor, gamma: str = "sin2",
                      a: float = 1.0, want_b: bool = False):
        """(loss [1] fp64, grad [n_weights] fp32, b [2N,3] or None): StandardVelocityLoss for given draws t [N] / [N,1]
        (one value per molecule repeated over its atoms) and z [N,3], and d loss / d weights, all on the device."""
        if weights.dtype != torch.float32 or weights.numel() != self.n_weights or not weights.is_contiguous():
            raise ValueError(f"weights must be a contiguous fp32 vector of {self.n_weights} elements")
        N = tb.pb.n_nodes
        t = t.to(self.device, torch.float32).reshape(-1).contiguous()
        z = z.to(self.device, torch.float32).contiguous()
        if t.numel() != N or tuple(z.shape) != (N, 3):
            raise ValueError("t must have one entry per atom and z must be [N,3]")
        loss = torch.empty(1, dtype=torch.float64, device=self.device)
        grad = torch.empty(self.n_weights, dtype=torch.float32, device=self.device)
        out_b = torch.empty(2 * N, 3, dtype=torch.float32, device=self.device) if want_b else None
        pb = tb.pb
        cb = _lib.TrainBatch(n_mol=pb.n_mol, n_nodes=N, n_edges=pb.n_edges, mol_ptr=pb.mol_ptr.data_ptr(),
                             edge_ptr=pb.edge_ptr.data_ptr(), atom_id=pb.atom_id.data_ptr(), edge_type=pb.edge_type.data_ptr(),
                             temp0=pb.temp0.data_ptr(), temp1=pb.temp1.data_ptr(), x0=tb.x0.data_ptr(), x1=tb.x1.data_ptr(),
                             t=t.data_ptr(), z=z.data_ptr())
        ip = _lib.Interpolant(gamma_kind=_lib.GAMMAS[gamma], a=float(a))
        wp, wn = self._workspace(tb)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tib_train_loss_grad(C.byref(self.desc), weights.data_ptr(), C.byref(cb), C.byref(ip),
                                                    loss.data_ptr(), grad.data_ptr(), out_b.data_ptr() if want_b else None,
                                                    wp, wn, self._stream()), "tib_train_loss_grad")
        return loss, grad, out_b

    def status(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tib_train_status(self._stream()), "tib_train_status")

    def adam_step(self, weights, grad, m, v, step: int, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                  max_grad_norm=1.0) -> torch.Tensor:
        """clip_grad_norm_(max_grad_norm) + torch.optim.Adam in place on the flat vectors; returns the device scalar
        |grad|_2^2 (before clipping)."""
        for name, x in (("weights", weights), ("grad", grad), ("m", m), ("v", v)):
            if x.dtype != torch.float32 or not x.is_contiguous() or x.numel() != weights.numel() or x.device != weights.device:
                raise ValueError(f"{name} must be a contiguous fp32 vector like weights")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tib_adam_step(weights.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), weights.numel(),
                                              int(step), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                              float(weight_decay), float(max_grad_norm if max_grad_norm else 0.0),
                                              self._scratch.data_ptr(), self._stream()), "tib_adam_step")
        return self._scratch


def gemm_f16x3(A, B, *, trans_a=False, trans_b=False, idx_a=None, idx_b=None, scale_a=1.0, scale_b=1.0, amax_a=None,
               out=None, c_idx=None, bias=None, mode=_lib.GEMM_STORE, split_k=False, M=None, N=None, K=None):
    """C[M][N] = sum_k A(m,k) B(n,k) through tib_gemm_f16x3 (tests / benchmarks).  A, B: 2-D fp32 CUDA tensors (row-major,
    may be column slices of wider tensors); see include/tib.h for the operand conventions."""
    lib = _lib.load()
    rows_a = idx_a.numel() if idx_a is not None else A.shape[0]
    rows_b = idx_b.numel() if idx_b is not None else B.shape[0]
    if M is None:
        M = A.shape[1] if trans_a else rows_a
    if N is None:
        N = B.shape[1] if trans_b else rows_b
    if K is None:
        K = rows_a if trans_a else A.shape[1]
    if out is None:
        out = torch.zeros(M, N, dtype=torch.float32, device=A.device)
    ptr = lambda x: x.data_ptr() if x is not None else None  # noqa: E731
    with torch.cuda.device(A.device):
        _lib.check(lib.tib_gemm_f16x3(M, N, K, A.data_ptr(), A.stride(0), int(trans_a), ptr(idx_a), float(scale_a), ptr(amax_a),
                                      B.data_ptr(), B.stride(0), int(trans_b), ptr(idx_b), float(scale_b), out.data_ptr(),
                                      out.stride(0), ptr(c_idx), ptr(bias), int(mode), int(split_k),
                                      C.c_void_p(torch.cuda.current_stream(A.device).cuda_stream)), "tib_gemm_f16x3")
    return out
