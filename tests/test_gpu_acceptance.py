"""GPU tier: the statistical acceptance check of BASELINE.json's north_star - "matching free-energy differences or TICA
histograms within statistical error" - between samples of the CUDA path and samples of the CPU oracle (the reference
arithmetic) drawn from the same inputs: 2048 conformers through the same flow, then the reference's post-processing chain
(z-matrix torsions -> (cos, sin) features -> TICA projection -> density histograms; reweighting weights -> ESS, dF with a
bootstrap interval; IQR outlier mask)."""
import numpy as np
import pytest
import torch

from tests._util import oracle_hp_sd, perturb_
from thermodynamic_interpolation_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N_MOL, N_ATOMS, STEPS = 2048, 9, 10


@pytest.fixture(scope="module")
def samples():
    """(x0, CUDA samples, oracle samples) [N_MOL, 9, 3]: fixed-grid Euler, tensor-core default mode vs the fp32 oracle."""
    from oracle import cpainn_oracle as co
    from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    torch.manual_seed(3)
    model = perturb_(cPaiNN(n_features=128, score_layers=2, temp_length=100), 4).eval()
    mb = synthetic_ambient_batch(N_MOL, N_ATOMS, T0=1000.0, T1=300.0, sigma=0.3, seed=11)
    hp, sd = oracle_hp_sd(model)
    with torch.no_grad():
        ref, _, _ = co.rollout(sd, hp, mb.x0, mb.atoms, mb.edge_index, mb.edge_type, mb.ptr.tolist(), method="euler",
                               n_step=STEPS + 1, T0=mb.T0, T1=mb.T1)
    shape = (N_MOL, N_ATOMS, 3)
    x0 = mb.x0.reshape(shape).numpy().copy()
    integ = MoleculeIntegrator(model.to(DEV), method="euler", n_step=STEPS + 1, save_frames=False)
    xts, _, _, _ = integ.rollout(mb.to(DEV))          # MolBatch.to moves the batch in place
    return x0, xts.reshape(shape), ref[-1].reshape(shape).numpy()


def _chain_refs(n):
    """A simple valid z-matrix: atom a is placed relative to (a-1, a-2, a-3); unused slots of rows 0..2 hold other atoms."""
    ref = []
    for a in range(n):
        prev = [a - 1, a - 2, a - 3]
        fill = [j for j in range(n) if j != a and j not in prev]
        ref.append([p if p >= 0 else fill.pop() for p in prev])
    return ref, list(range(n))


def test_tica_histograms_match_within_bootstrap_error(samples):
    from oracle import analysis_oracle as ao
    from oracle import zmatrix_oracle as zo
    from thermodynamic_interpolation_b200 import analysis as A
    x0, x_cuda, x_ref = samples
    ref_atoms, order = _chain_refs(N_ATOMS)
    # reference chain on the oracle samples (numpy)
    z_ref = zo.construct_z_matrix_batch(x_ref, ref_atoms, order)
    tors_ref = z_ref[:, 2:, 2]
    feats_ref = ao.enc(tors_ref)
    mean, R, lam = ao.tica_fit(feats_ref, lagtime=1, dim=2)
    proj_ref = ao.tica_transform(feats_ref, mean, R)
    scale = np.abs(proj_ref).max(0)                     # histogram window per component (synthetic data has no fixed scale)
    bins = 40
    # product chain on the CUDA samples: z-matrix kernel -> projection + histogram kernel, the torsions read in place
    z = A.construct_z_matrix_batch(x_cuda, ref_atoms, order)
    err_t = np.abs(np.angle(np.exp(1j * (A.gen_torsions(z).cpu().numpy() - tors_ref)))).max()
    assert err_t < 5e-3, err_t                          # torsion angles of the two sample sets (atan2 amplifies near-planar cases)
    m1, R1, lam1 = A.tica_fit(torch.from_numpy(ao.enc(A.gen_torsions(z).cpu().numpy())).to(DEV), lagtime=1, dim=2)
    np.testing.assert_allclose(lam1.cpu().numpy(), lam, rtol=1e-3, atol=1e-4)
    for k in range(2):
        lo, hi = -1.05 * scale[k], 1.05 * scale[k]
        Rk = torch.from_numpy(np.ascontiguousarray(R[:, k:k + 1]))
        proj, dens = A.tica_project(z.reshape(N_MOL, -1), torch.from_numpy(mean), Rk, bins=bins, lo=lo, hi=hi,
                                    col0=3 * 2 + 2, col_step=3, n_tors=N_ATOMS - 3)
        h_ref = ao.density_hist(proj_ref[:, k], bins=bins, lo=lo, hi=hi)
        # statistical error of the reference histogram: bootstrap over conformers
        rng = np.random.default_rng(0)
        boots = np.stack([ao.density_hist(proj_ref[rng.integers(0, N_MOL, N_MOL), k], bins=bins, lo=lo, hi=hi) for _ in range(200)])
        sigma = boots.std(0)
        diff = np.abs(dens[0].cpu().numpy() - h_ref)
        print(f"[accept] tIC{k + 1}: eigenvalue {lam[k]:.4f}, max |rho_cuda - rho_ref| = {diff.max():.3e}, median bootstrap sigma {np.median(sigma):.3e}, "
              f"max |proj_cuda - proj_ref| = {np.abs(proj[:, 0].cpu().numpy() - proj_ref[:, k]).max():.3e}")
        assert (diff <= 3.0 * sigma + 1e-9).all()


def test_free_energy_and_ess_match_within_bootstrap_error(samples):
    from thermodynamic_interpolation_b200 import analysis as A, stats as S
    x0, x_cuda, x_ref = samples
    # synthetic reduced energies: harmonic, E = 1/2 |x|^2 (1000 / T)
    E0 = 0.5 * (x0.reshape(N_MOL, -1) ** 2).sum(1).astype(np.float64)
    e1 = lambda x: 0.5 * (np.asarray(x, dtype=np.float64).reshape(N_MOL, -1) ** 2).sum(1) * (1000.0 / 300.0)  # noqa: E731
    E1_ref, E1_cuda = e1(x_ref), e1(x_cuda.cpu().numpy())
    shift = float((E1_ref - E0).min())
    nd = np.zeros(N_MOL)
    dF_ref, ci = A.bootstrap_dF(E0, E1_ref - shift, nd, n_bootstrap=300, k=100, seed=7)
    dF_cuda, _ = A.bootstrap_dF(E0, E1_cuda - shift, nd, n_bootstrap=2, k=100, seed=7)
    print(f"[accept] dF oracle samples {dF_ref:.6f} in [{ci[0]:.6f}, {ci[1]:.6f}], CUDA samples {dF_cuda:.6f}")
    assert ci[0] <= dF_cuda <= ci[1]
    # device-side statistics of the CUDA samples agree with the host formulae on the same numbers
    dev = torch.device(DEV)
    tot = S.finalize(S.reweight_partials(torch.from_numpy(E0).to(dev), torch.from_numpy(E1_cuda - shift).to(dev)).cpu())
    w = np.exp(-(E1_cuda - shift - E0))
    np.testing.assert_allclose(tot["ess"], w.sum() ** 2 / (w ** 2).sum(), rtol=1e-10)
    w_ref = np.exp(-(E1_ref - shift - E0))
    ess_ref = w_ref.sum() ** 2 / (w_ref ** 2).sum()
    print(f"[accept] ESS oracle samples {ess_ref:.3f}, CUDA samples {tot['ess']:.3f}")
    assert abs(tot["ess"] - ess_ref) / ess_ref < 1e-2
    # the IQR mask with global percentiles keeps the same conformers
    keep_cuda = A.filter_iqr(torch.from_numpy(w).to(dev), k=100).cpu().numpy()
    q75, q25 = np.percentile(w_ref, [75, 25])
    keep_ref = (w_ref > q25 - 100 * (q75 - q25)) & (w_ref < q75 + 100 * (q75 - q25))
    assert (keep_cuda != keep_ref).sum() <= 2
