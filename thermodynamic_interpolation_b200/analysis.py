"""Drop-in for the post-processing step of the reference's analysis that runs over every sample:
`mdqm9/analysis/utils/z_matrix.py::construct_z_matrix_batch` and the column selectors of
`mdqm9/analysis/results_00031.py:140-149`.  The atom ordering / reference triplets come from the reference's
RDKit graph walk (`sort_atoms.compute_atom_order_and_references_groups`, host side, once per molecule) and are
passed in; the per-conformer arithmetic (distances, acos angles, atan2 torsions) is one CUDA kernel
(csrc/postproc.cuh) over samples that are already on the device after the rollout.  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib


def construct_z_matrix_batch(X_batch: torch.Tensor, ref_atoms: Sequence[Sequence[int]],
                             placing_order: Optional[Sequence[int]] = None) -> torch.Tensor:
    """X_batch [n_conf, n_atoms, 3] (CUDA, fp32) -> z_matrix [n_conf, n_atoms-1, 3] = (distance, angle, torsion),
    same indexing as the reference (z_matrix.py:56-102): row a-1 describes placed atom a relative to
    ref_atoms[a] = (distance partner, angle partner, torsion partner)."""
    if X_batch.device.type != "cuda":
        raise RuntimeError("thermodynamic_interpolation_b200 runs on CUDA devices only (no CPU fallback)")
    if X_batch.dim() != 3 or X_batch.shape[-1] != 3:
        raise ValueError("X_batch must be [n_conformations, n_atoms, 3]")
    n_conf, n_atoms, _ = X_batch.shape
    if len(ref_atoms) != n_atoms or any(len(t) != 3 for t in ref_atoms):
        raise ValueError("ref_atoms must hold one (distance, angle, torsion) reference triplet per atom")
    if placing_order is None:
        placing_order = list(range(n_atoms))
    if len(placing_order) != n_atoms:
        raise ValueError("placing_order must list every atom once")
    idx = torch.tensor([list(t) for t in ref_atoms], dtype=torch.int32)
    order = torch.tensor(list(placing_order), dtype=torch.int32)
    if int(idx.min()) < 0 or int(idx.max()) >= n_atoms or int(order.min()) < 0 or int(order.max()) >= n_atoms:
        raise ValueError("atom index out of range")
    x = X_batch.to(torch.float32).contiguous()
    z = torch.empty((n_conf, n_atoms - 1, 3), dtype=torch.float32, device=x.device)
    idx, order = idx.to(x.device), order.to(x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.tib_zmatrix(x.data_ptr(), n_conf, n_atoms, order.data_ptr(), idx.data_ptr(), z.data_ptr(),
                                   C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "tib_zmatrix")
    return z


def gen_torsions(z_matrix: torch.Tensor) -> torch.Tensor:       # results_00031.py:140-141
    return z_matrix[:, 2:, 2]


def gen_bond_angles(z_matrix: torch.Tensor) -> torch.Tensor:    # results_00031.py:144-145
    return z_matrix[:, 1:, 1]


def gen_bond_lengths(z_matrix: torch.Tensor) -> torch.Tensor:   # results_00031.py:148-149
    return z_matrix[:, :, 0]
