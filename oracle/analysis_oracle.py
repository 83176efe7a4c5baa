"""TEST INFRASTRUCTURE ONLY (oracle/).  numpy / scipy restatement of the reference's post-processing that follows the
z-matrix: the torsion encoding and TICA projection of mdqm9/plots/10506_main.ipynb cell 3, its density histograms (cell 4),
and the bootstrap of results_00031.py:29-45.

PARITY STATUS: the bootstrap / free-energy part is pinned (tests/golden/stats.npz holds the outputs of the reference's own
calc_phis_tfep / calc_tfep_dF / filter_iqr inside the loop of results_00031.py:29-45).  TICA is `deeptime.decomposition.TICA`
(deeptime is an un-vendored dependency, absent here): restated from its published algorithm - PARITY UNPINNED; this
restatement solves the generalised eigenproblem directly (scipy.linalg.eigh(C0t, C00)), independently of the whitening route
the product takes."""
import numpy as np
import scipy.linalg


def enc(x):
    """10506_main.ipynb cell 3: (cos, sin) pairs per torsion, merged along the last axis."""
    e = np.stack((np.cos(x), np.sin(x)), axis=-1)
    return e.reshape(e.shape[0], -1)


def tica_fit(features, lagtime, dim=2):
    """Reversible TICA with the mean removed and kinetic-map scaling: returns (mean, R [d, dim], eigenvalues)."""
    f = np.asarray(features, dtype=np.float64)
    X, Y = f[:-lagtime], f[lagtime:]
    n = X.shape[0]
    mean = 0.5 * (X.mean(0) + Y.mean(0))
    Xc, Yc = X - mean, Y - mean
    c00 = (Xc.T @ Xc + Yc.T @ Yc) / (2.0 * n)
    c0t = (Xc.T @ Yc + Yc.T @ Xc) / (2.0 * n)
    lam, v = scipy.linalg.eigh(c0t, c00)           # v' C00 v = I
    order = np.argsort(lam)[::-1][:dim]
    return mean, v[:, order] * lam[order], lam[order]


def tica_transform(features, mean, R):
    return (np.asarray(features, dtype=np.float64) - mean) @ R


def density_hist(values, bins=80, lo=-2.5, hi=2.5, weights=None):
    h, _ = np.histogram(values, bins=np.linspace(lo, hi, bins + 1), weights=weights, density=True)
    return h


def bootstrap_dF(E0s, E1s, nd, calc_phis_tfep, calc_tfep_dF, n_bootstrap, k, seed):
    """The loop of gen_free_energy_tfep_md_ti (results_00031.py:29-45) around the given free-energy functions."""
    np.random.seed(seed)
    phis, _ = calc_phis_tfep(E0s=E0s, E1s=E1s, neg_dlogps_ti=nd, k=k)
    est = np.zeros(n_bootstrap)
    for i in range(n_bootstrap):
        idx = np.random.choice(np.arange(len(phis)), len(phis), replace=True)
        pb, _ = calc_phis_tfep(E0s=E0s[idx], E1s=E1s[idx], neg_dlogps_ti=nd[idx], k=k)
        est[i] = calc_tfep_dF(phis=pb, weights=np.ones_like(pb))
    return calc_tfep_dF(phis=phis, weights=np.ones_like(phis)), [np.percentile(est, 2.5), np.percentile(est, 97.5)]
