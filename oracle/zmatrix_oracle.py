"""TEST INFRASTRUCTURE ONLY (oracle/).  numpy restatement of
mdqm9/analysis/utils/z_matrix.py::construct_z_matrix_batch (:56-102) with compute_distance / compute_angle /
compute_torsion of mdqm9/analysis/utils/mol_geometry.py:25-81.  Pinned against the unmodified reference by
tests/golden/zmatrix.npz (oracle/make_golden.py::zmatrix_case)."""
from __future__ import annotations

import numpy as np


def construct_z_matrix_batch(X, ref_atoms, placing_order=None):
    X = np.asarray(X, dtype=np.float32)
    n_atoms = X.shape[1]
    if placing_order is None:
        placing_order = list(range(len(ref_atoms)))
    i3 = [t[0] for t in ref_atoms]
    i2 = [t[1] for t in ref_atoms]
    i1 = [t[2] for t in ref_atoms]
    x4 = X[:, placing_order, :]
    x3 = X[:, i3[1:], :]
    x2 = X[:, i2[2:], :]
    x1 = X[:, i1[3:], :]
    z = np.zeros((X.shape[0], n_atoms - 1, 3), dtype=np.float32)
    z[:, :, 0] = np.linalg.norm(x3 - x4[:, 1:, :], axis=-1)                        # mol_geometry.py:25-37
    u, w = x4[:, 2:, :] - x3[:, 1:, :], x2 - x3[:, 1:, :]                          # mol_geometry.py:40-56
    z[:, 1:, 1] = np.arccos((u * w).sum(-1) / (np.linalg.norm(u, axis=-1) * np.linalg.norm(w, axis=-1)))
    a, b, c, d = x1, x2[:, 1:, :], x3[:, 2:, :], x4[:, 3:, :]                      # mol_geometry.py:59-81
    x12, x23, x34 = b - a, c - b, d - c
    c2334 = np.cross(x23, x34)
    y = np.linalg.norm(x23, axis=-1) * (x12 * c2334).sum(-1)
    x = (np.cross(x12, x23) * c2334).sum(-1)
    z[:, 2:, 2] = np.arctan2(y, x)
    return z
