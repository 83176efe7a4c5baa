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


# ---- TICA of the torsion features, histograms, outlier filter, bootstrap ----------------------------
def tica_fit(features: torch.Tensor, lagtime: int, dim: int = 2, epsilon: float = 1e-6):
    """`deeptime.decomposition.TICA(lagtime, dim).fit(features)` restated (plots/10506_main.ipynb cell 3; deeptime 0.4 is an
    un-vendored pin of the reference and absent here: its published algorithm, "parity unpinned").  features [T, d] is ONE
    time series.  Reversible covariance estimate with the data mean removed (deeptime Covariance(remove_data_mean=True,
    reversible=True, bessels_correction=False)): X = f[:-lag], Y = f[lag:], mean = (mean X + mean Y) / 2,
    C00 = (Xc'Xc + Yc'Yc) / 2n, C0t = (Xc'Yc + Yc'Xc) / 2n; whitening L = V diag(s^-1/2) over the eigenvalues s > epsilon of C00
    (spd_inv_split); eigen-decomposition of L' C0t L, descending; R = L W scaled by the eigenvalues (scaling='kinetic_map').
    Returns (mean [d], R [d, dim], eigenvalues [dim]) as fp64 tensors on the device of `features`."""
    f = features.to(torch.float64)
    if f.dim() != 2 or f.shape[0] <= lagtime:
        raise ValueError("features must be [T, d] with T > lagtime")
    X, Y = f[:-lagtime], f[lagtime:]
    n = X.shape[0]
    mean = 0.5 * (X.mean(0) + Y.mean(0))
    Xc, Yc = X - mean, Y - mean
    c00 = (Xc.T @ Xc + Yc.T @ Yc) / (2.0 * n)
    c0t = (Xc.T @ Yc + Yc.T @ Xc) / (2.0 * n)
    s, V = torch.linalg.eigh(c00)
    keep = s > epsilon
    L = V[:, keep] / torch.sqrt(s[keep])
    # canonical signs (deeptime spd_inv_split(canonical_signs=True)): the largest-magnitude entry of every column is positive
    idx = L.abs().argmax(0)
    L = L * torch.sign(L[idx, torch.arange(L.shape[1], device=L.device)])
    lam, W = torch.linalg.eigh(L.T @ c0t @ L)
    order = torch.argsort(lam, descending=True)[:dim]
    lam, W = lam[order], W[:, order]
    R = (L @ W) * lam
    return mean.contiguous(), R.contiguous(), lam


def tica_project(torsions: torch.Tensor, mean: torch.Tensor, R: torch.Tensor, *, weights: Optional[torch.Tensor] = None,
                 bins: int = 80, lo: float = -2.5, hi: float = 2.5, col0: int = 0, col_step: int = 1,
                 n_tors: Optional[int] = None):
    """enc() + TICA.transform + the density histograms of plots/10506_main.ipynb cells 3-4 in one pass over the samples
    (csrc/postproc.cuh k_tica_project).  torsions: CUDA fp32 [n_conf, n_cols]; feature j reads column col0 + j * col_step
    (a flattened z-matrix can be read in place).  Returns (proj [n_conf, dim] fp32, density [dim, bins] fp64 normalised like
    plt.hist(density=True): sum(density) * bin_width = 1 over the samples that fall inside [lo, hi])."""
    if torsions.device.type != "cuda":
        raise RuntimeError("thermodynamic_interpolation_b200 runs on CUDA devices only (no CPU fallback)")
    t = torsions.to(torch.float32).contiguous()
    n_conf, n_cols = t.shape
    if n_tors is None:
        n_tors = (n_cols - col0 + col_step - 1) // col_step
    dim = int(R.shape[1])
    if tuple(mean.shape) != (2 * n_tors,) or int(R.shape[0]) != 2 * n_tors:
        raise ValueError("mean / R do not match 2 * n_tors features")
    dev = t.device
    mean, R = mean.to(dev, torch.float64).contiguous(), R.to(dev, torch.float64).contiguous()
    w = None if weights is None else weights.to(dev, torch.float64).contiguous()
    proj = torch.empty((n_conf, dim), dtype=torch.float32, device=dev)
    hist = torch.zeros((dim, bins), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tib_tica_project(t.data_ptr(), n_conf, n_tors, n_cols, col0, col_step, mean.data_ptr(), R.data_ptr(),
                                                dim, None if w is None else w.data_ptr(), proj.data_ptr(), hist.data_ptr(), bins,
                                                float(lo), float(hi), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "tib_tica_project")
    width = (hi - lo) / bins
    density = hist / (hist.sum(1, keepdim=True).clamp_min(1e-300) * width)
    return proj, density


def filter_iqr(x: torch.Tensor, k: Optional[float] = 10, group=None) -> torch.Tensor:
    """sensititvity.filter_iqr (mdqm9/analysis/utils/sensititvity.py:4-12) with GLOBAL percentiles: across ranks the 1-D
    vector is all-gathered first, the returned mask is for the LOCAL elements."""
    if k is None:
        return torch.ones(x.shape, dtype=torch.bool, device=x.device)
    from . import dist as D
    full = D.gather_samples(x, group) if D._world(group) > 1 else x
    q = torch.quantile(full.to(torch.float64), torch.tensor([0.25, 0.75], dtype=torch.float64, device=x.device))   # linear interpolation = np.percentile
    q25, q75 = float(q[0]), float(q[1])
    iqr = q75 - q25
    return (x > q25 - k * iqr) & (x < q75 + k * iqr)


def bootstrap_dF(E0s, E1s, neg_dlogps, n_bootstrap: int = 1000, k: Optional[float] = None, seed: Optional[int] = None):
    """gen_free_energy_tfep_md_ti (mdqm9/analysis/results_00031.py:29-45): dF = -log mean exp(-phi) over the IQR-kept
    samples and its 95 % bootstrap interval.  Host side (numpy), as in the reference; `seed` = np.random.seed."""
    import numpy as np
    E0s, E1s, nd = (np.asarray(a, dtype=np.float64) for a in (E0s, E1s, neg_dlogps))

    def phis_of(e0, e1, d):                                # free_energy.calc_phis_tfep (free_energy.py:25-38)
        phi = e1 - e0 + d
        if k is not None:
            w = np.exp(-phi)
            q75, q25 = np.percentile(w, [75, 25])
            keep = (w > q25 - k * (q75 - q25)) & (w < q75 + k * (q75 - q25))
            phi = -np.log(w[keep])                         # the reference round-trips through exp / log
        return phi

    def dF_of(phi):                                        # free_energy.calc_tfep_dF (free_energy.py:41-46)
        return -np.log(np.mean(np.exp(-phi)))

    if seed is not None:
        np.random.seed(seed)
    phis = phis_of(E0s, E1s, nd)
    est = np.zeros(n_bootstrap)
    n = len(phis)              # the reference resamples len(phis) indices - the count AFTER the filter - from the unfiltered arrays
    for i in range(n_bootstrap):
        idx = np.random.choice(np.arange(n), n, replace=True)
        est[i] = dF_of(phis_of(E0s[idx], E1s[idx], nd[idx]))
    return dF_of(phis), [np.percentile(est, 2.5), np.percentile(est, 97.5)]
