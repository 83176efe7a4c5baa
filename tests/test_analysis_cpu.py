"""CPU tier: host-side post-processing (bootstrap of the free-energy difference) against the reference's frozen outputs, and
the TICA restatements against each other on a synthetic metastable time series."""
import numpy as np
import torch

from oracle import analysis_oracle as ao
from tests._util import load_golden


def test_bootstrap_dF_matches_reference():
    """analysis.bootstrap_dF == the loop of results_00031.py:29-45 around the reference's own free-energy functions."""
    from thermodynamic_interpolation_b200.analysis import bootstrap_dF
    g = load_golden("stats")
    for tag, k in (("none", None), ("k3", 3)):
        dF, ci = bootstrap_dF(g["E0"], g["E1"], g["neg_dlogp"], n_bootstrap=200, k=k, seed=123)
        np.testing.assert_allclose(dF, g[f"boot_dF_{tag}"], rtol=1e-12)
        np.testing.assert_allclose(ci, g[f"boot_ci_{tag}"], rtol=1e-12)


def _series(n=6000, seed=5):
    """Torsion-like time series with one slow two-state process and fast noise."""
    rng = np.random.default_rng(seed)
    state = np.zeros(n, dtype=int)
    for t in range(1, n):
        state[t] = state[t - 1] ^ (rng.random() < 0.01)
    tors = np.stack([np.where(state, 1.0, -1.2) + 0.3 * rng.normal(size=n), rng.uniform(-np.pi, np.pi, n),
                     0.5 * np.where(state, -1.0, 0.7) + 0.4 * rng.normal(size=n)], axis=1)
    return tors.astype(np.float32)


def test_tica_fit_two_routes_agree():
    """analysis.tica_fit (whitening + symmetric eigenproblem, torch) and the oracle (generalised eigenproblem, scipy) give the
    same eigenvalues and, up to the sign of each component, the same projection; the slow process is the first component."""
    from thermodynamic_interpolation_b200.analysis import tica_fit
    tors = _series()
    feats = ao.enc(tors)
    for lag in (1, 8):
        m0, R0, lam0 = ao.tica_fit(feats, lag, dim=2)
        m1, R1, lam1 = tica_fit(torch.from_numpy(feats), lag, dim=2)
        np.testing.assert_allclose(lam1.numpy(), lam0, rtol=1e-8)
        np.testing.assert_allclose(m1.numpy(), m0, rtol=1e-12, atol=1e-14)
        p0, p1 = ao.tica_transform(feats, m0, R0), ao.tica_transform(feats, m1.numpy(), R1.numpy())
        for k in range(2):
            sign = np.sign((p0[:, k] * p1[:, k]).sum())
            np.testing.assert_allclose(sign * p1[:, k], p0[:, k], rtol=1e-6, atol=1e-8)
        assert lam0[0] > 0.8 and lam0[1] < 0.5
