"""GPU tier: internal coordinates of sampled conformers (tib_zmatrix) against the unmodified
reference's construct_z_matrix_batch (tests/golden/zmatrix.npz) and the numpy oracle at production size."""
import numpy as np
import pytest
import torch

from tests._util import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _wrap(a, b):
    d = np.abs(a - b)
    return np.minimum(d, 2 * np.pi - d)          # torsions live on a circle: -pi and +pi are the same angle


@pytest.mark.parametrize("tag", ["a9", "a25"])
def test_zmatrix_matches_reference_golden(tag):
    from thermodynamic_interpolation_b200.analysis import construct_z_matrix_batch, gen_bond_angles, gen_bond_lengths, gen_torsions
    g = load_golden("zmatrix")
    X = torch.from_numpy(g[f"{tag}::X"]).to(DEV)
    z = construct_z_matrix_batch(X, g[f"{tag}::ref"].tolist(), g[f"{tag}::order"].tolist())
    ref = g[f"{tag}::z"]
    zc = z.cpu().numpy()
    assert zc.shape == ref.shape
    np.testing.assert_allclose(zc[..., :2], ref[..., :2], atol=5e-6, rtol=0)
    assert _wrap(zc[..., 2], ref[..., 2]).max() < 5e-6
    assert (zc[:, 0, 1:] == 0).all() and (zc[:, 1, 2] == 0).all()
    assert gen_torsions(z).shape == (ref.shape[0], ref.shape[1] - 2) and gen_bond_angles(z).shape[1] == ref.shape[1] - 1
    assert gen_bond_lengths(z).shape == ref.shape[:2]


def test_zmatrix_at_production_size_and_invariance():
    """1e6 conformers x 9 atoms (BASELINE cfg 4 scale): equals the oracle on a slice and is invariant under rigid motion."""
    from oracle import zmatrix_oracle as zo
    from thermodynamic_interpolation_b200.analysis import construct_z_matrix_batch
    g = load_golden("zmatrix")
    ref, order = g["a9::ref"].tolist(), g["a9::order"].tolist()
    gen = torch.Generator(device=DEV).manual_seed(7)
    X = torch.randn(1_000_000, 9, 3, device=DEV, generator=gen)
    z = construct_z_matrix_batch(X, ref, order)
    assert torch.isfinite(z).all()
    sl = slice(123_456, 123_456 + 512)
    zo_ = zo.construct_z_matrix_batch(X[sl].cpu().numpy(), ref, order)
    zc = z[sl].cpu().numpy()
    np.testing.assert_allclose(zc[..., :2], zo_[..., :2], atol=1e-5, rtol=0)
    assert _wrap(zc[..., 2], zo_[..., 2]).max() < 1e-4            # atan2 near +-pi / small arguments amplifies ulps
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    R = q.to(torch.float32).to(DEV)
    z2 = construct_z_matrix_batch(X[:4096] @ R.T + 0.5, ref, order)
    d = (z2 - z[:4096]).abs().cpu().numpy()
    assert d[..., :2].max() < 2e-4 and np.minimum(d[..., 2], 2 * np.pi - d[..., 2]).max() < 2e-3


def test_zmatrix_argument_errors():
    from thermodynamic_interpolation_b200.analysis import construct_z_matrix_batch
    X = torch.zeros(4, 5, 3, device=DEV)
    with pytest.raises(ValueError):
        construct_z_matrix_batch(X, [[0, 1, 2]] * 4)
    with pytest.raises(ValueError):
        construct_z_matrix_batch(X, [[0, 1, 7]] * 5)
    with pytest.raises(RuntimeError, match="CUDA"):
        construct_z_matrix_batch(X.cpu(), [[0, 1, 2]] * 5)
