"""GPU tier: the exact divergence / dlogp path (`return_dlogp=True`).

tib_drift_div propagates 3 * max_atoms forward-mode tangent directions through dual-number CUDA kernels
(csrc/simt_tangent.cuh); the reference gets the same number from 3n autograd passes
(ODEWrapper.compute_divergence, ambient ode_wrapper.py:59-91).  Checked against
 (1) the reference's own outputs frozen in tests/golden/ambient_f32.npz (divergence and a dlogp rollout),
 (2) the CPU oracle's autograd divergence on seeded batches - mixed molecule sizes, both variants, every
     feature width the kernels are built for (D = 3 directions per pass, and D = 1 for F = 256),
 (3) invariance of the divergence under rotations and translations at BASELINE cfg-2 shape."""
import numpy as np
import pytest
import torch

from tests._util import golden_batch, golden_model, load_golden, oracle_hp_sd, oracle_temps, perturb_

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    assert torch.cuda.is_available(), "the gpu tier needs a CUDA device"
    from thermodynamic_interpolation_b200 import _lib
    _lib.load()
    yield


def _close(a, b, rtol, atol_rel, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    err = np.abs(a - b).max() / scale
    print(f"[parity] {what}: max|diff|/max|ref| = {err:.3e}")
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol_rel * scale)


def _oracle_div(model, batch, x, t):
    from oracle import cpainn_oracle as co
    hp, sd = oracle_hp_sd(model)
    atoms = batch.atoms if hp.variant == "ambient" else batch.atom_number
    cpu = lambda v: v.detach().cpu()  # noqa: E731
    temps = {k: cpu(v) for k, v in oracle_temps(batch, hp).items()}
    return co.divergence(sd, hp, cpu(x), t, cpu(atoms), cpu(batch.edge_index), cpu(batch.edge_type),
                         batch.ptr.tolist(), **temps)


def test_divergence_matches_reference_golden():
    """ODEWrapper(return_dlogp=True).forward -> (b, -div * 1e-2) against the unmodified reference."""
    from thermodynamic_interpolation_b200.ambient.models.ode_wrapper import ODEWrapper
    from thermodynamic_interpolation_b200 import _lib
    g = load_golden("ambient_f32")
    model = golden_model(g, DEV)
    batch = golden_batch(g).to(DEV)
    x0 = batch.x0.clone()
    b, negdiv = ODEWrapper(model, return_dlogp=True)(torch.tensor(0.37), (x0, torch.zeros(int(g["n_mol"]), device=DEV)), batch, [0])
    _close((-negdiv).cpu().numpy(), g["div_t037_scaled"], rtol=2e-4, atol_rel=2e-5, what="divergence*1e-2 vs reference")
    # the drift that comes with it is the fp32 CUDA-core drift (same arithmetic, smaller edge tiles, so the
    # per-node sums over incoming edges are grouped differently) and matches the reference's
    model.set_math(_lib.MATH_FP32_SIMT)
    b_simt = ODEWrapper(model)(torch.tensor(0.37), x0, batch, [0])
    _close(b.cpu().numpy(), b_simt.cpu().numpy(), rtol=1e-4, atol_rel=2e-6, what="drift of the tangent path vs fp32 path")
    _close(b.cpu().numpy(), g["drift"][1], rtol=1e-4, atol_rel=1e-5, what="drift of the tangent path vs reference")
    bm, rm = ODEWrapper(model, return_dlogp=True, reverse_ode=True)(torch.tensor(0.37), (x0, None), batch, [0])
    assert torch.equal(bm, -b) and torch.equal(rm, -negdiv)


def test_dlogp_rollout_matches_reference_golden():
    """MoleculeIntegrator(return_dlogp=True).rollout == the reference's (xts, dlogp * 1e2, nfe, batch)."""
    from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    g = load_golden("ambient_f32")
    model = golden_model(g, DEV)
    batch = golden_batch(g).to(DEV)
    xts, dlogp, nfe, bvec = MoleculeIntegrator(model, method="euler", n_step=3, return_dlogp=True).rollout(batch)
    assert nfe == 2 and tuple(dlogp.shape) == (3, int(g["n_mol"])) and torch.equal(bvec, batch.batch)
    _close(xts.cpu().numpy(), g["euler_dlogp_xts"], rtol=1e-4, atol_rel=2e-5, what="euler+dlogp frames vs reference")
    _close(dlogp.cpu().numpy(), g["euler_dlogp"], rtol=2e-4, atol_rel=2e-5, what="dlogp*1e2 vs reference")


CASES = [  # variant, F, L, molecule sizes
    ("ambient", 32, 3, [9, 9, 9]),
    ("ambient", 64, 2, [5, 3, 7, 2]),
    ("latent_multi", 128, 2, [4, 6, 3]),
    ("latent_single", 64, 2, [6, 6]),
    ("ambient", 256, 1, [3, 4]),
]


def _make(variant, F, L, sizes, seed=3):
    from thermodynamic_interpolation_b200 import batch as B
    torch.manual_seed(seed)
    if variant == "ambient":
        from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
        model = cPaiNN(n_features=F, score_layers=L, temp_length=100)
        batch = B.synthetic_ambient_batch(len(sizes), sizes, seed=seed + 1)
    else:
        from thermodynamic_interpolation_b200.latent.models.cpainn import cPaiNN
        multi = variant == "latent_multi"
        model = cPaiNN(n_features=F, score_layers=L, temp_length=75,
                       temperatures=[300, 400, 500, 600, 700, 800, 900, 1000] if multi else [300])
        batch = B.synthetic_latent_batch(len(sizes), sizes, T=800 if multi else None, seed=seed + 1)
    perturb_(model, seed + 2, 0.05)
    return model.eval().to(DEV), batch


@pytest.mark.parametrize("variant,F,L,sizes", CASES)
def test_divergence_matches_oracle_autograd(variant, F, L, sizes):
    model, batch = _make(variant, F, L, sizes)
    ref = _oracle_div(model, batch, batch.x0, 0.41).numpy()
    eng = model.engine()
    pb = eng.prepare(batch.clone().to(DEV))
    b, div = eng.drift_div(pb, batch.x0.to(DEV), 0.41)
    _close(div.cpu().numpy(), ref, rtol=5e-4, atol_rel=5e-5, what=f"divergence {variant} F={F} sizes={sizes} vs oracle autograd")


def test_latent_rollout_with_dlogp_and_reverse_matches_oracle():
    """Latent variant: unscaled dlogp, 3-tuple return; reverse_ode integrates (-b, +div) on the reversed grid."""
    from oracle import cpainn_oracle as co
    from thermodynamic_interpolation_b200.latent.integrators import MoleculeIntegrator
    model, batch = _make("latent_multi", 64, 2, [5, 5, 5])
    hp, sd = oracle_hp_sd(model)
    temps = oracle_temps(batch, hp)
    for reverse in (False, True):
        for method, n_step in (("euler", 4), ("rk4", 3)):
            xo, dlo, _ = co.rollout(sd, hp, batch.x0, batch.atom_number, batch.edge_index, batch.edge_type, batch.ptr.tolist(),
                                    method=method, n_step=n_step, return_dlogp=True, reverse_ode=reverse, **temps)
            xts, dlogp, bvec = MoleculeIntegrator(model, method=method, n_step=n_step, return_dlogp=True,
                                                  reverse_ode=reverse).rollout(batch.clone().to(DEV))
            _close(xts.cpu().numpy(), xo.numpy(), rtol=1e-4, atol_rel=2e-5, what=f"latent {method} reverse={reverse} frames")
            _close(dlogp.cpu().numpy(), dlo.numpy(), rtol=5e-4, atol_rel=5e-5, what=f"latent {method} reverse={reverse} dlogp")


@pytest.mark.parametrize("tol", [1e-6, 3e-7])
def test_dopri5_with_dlogp_matches_oracle(tol):
    """Tuple-state dopri5 inside libtib.so (tib_rollout_dopri5_dlogp: max of per-component RMS norms, Hairer initial step,
    0.9 / 0.2 / 10 controller, dense output) against the oracle's torchdiffeq restatement: same step sequence (equal NFE),
    frames and dlogp within fp32."""
    from oracle import cpainn_oracle as co
    from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    model, batch = _make("ambient", 32, 2, [4, 4, 4])
    hp, sd = oracle_hp_sd(model)
    stats = {}
    xo, dlo, nfe_o = co.rollout(sd, hp, batch.x0, batch.atoms, batch.edge_index, batch.edge_type, batch.ptr.tolist(),
                                method="dopri5", n_step=5, atol=tol, rtol=tol, return_dlogp=True, stats=stats,
                                **oracle_temps(batch, hp))
    integ = MoleculeIntegrator(model, method="dopri5", n_step=5, atol=tol, rtol=tol, return_dlogp=True)
    xts, dlogp, nfe, _ = integ.rollout(batch.clone().to(DEV))
    assert nfe == nfe_o, (nfe, nfe_o)
    _close(xts.cpu().numpy(), xo.numpy(), rtol=1e-4, atol_rel=2e-5, what=f"dopri5+dlogp frames vs oracle (tol {tol})")
    _close(dlogp.cpu().numpy()[1:], dlo.numpy()[1:], rtol=1e-3, atol_rel=1e-4, what=f"dopri5 dlogp*1e2 vs oracle (tol {tol})")


def test_dopri5_with_dlogp_loose_tolerance_is_a_valid_solution():
    """At rtol = atol = 1e-5 this problem takes three accepted steps and the third sits at an error ratio of 1.00 +- 0.02: the
    oracle (torch matmul over the stage axis) rejects it (ratio 1.003), the CUDA stage kernels (sequential fmaf) accept it
    (0.980) - both are dopri5 solutions at that tolerance.  The step sequences may differ by those attempts; the solutions agree
    to the global error of three steps."""
    from oracle import cpainn_oracle as co
    from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    model, batch = _make("ambient", 32, 2, [4, 4, 4])
    hp, sd = oracle_hp_sd(model)
    xo, dlo, nfe_o = co.rollout(sd, hp, batch.x0, batch.atoms, batch.edge_index, batch.edge_type, batch.ptr.tolist(),
                                method="dopri5", n_step=5, atol=1e-5, rtol=1e-5, return_dlogp=True, **oracle_temps(batch, hp))
    integ = MoleculeIntegrator(model, method="dopri5", n_step=5, atol=1e-5, rtol=1e-5, return_dlogp=True)
    xts, dlogp, nfe, _ = integ.rollout(batch.clone().to(DEV))
    assert abs(nfe - nfe_o) <= 12, (nfe, nfe_o)
    _close(xts.cpu().numpy(), xo.numpy(), rtol=1e-2, atol_rel=5e-3, what="dopri5+dlogp frames vs oracle (tol 1e-5)")
    _close(dlogp.cpu().numpy()[1:], dlo.numpy()[1:], rtol=5e-2, atol_rel=5e-2, what="dopri5 dlogp*1e2 vs oracle (tol 1e-5)")


def test_divergence_is_invariant_under_rigid_motion_at_cfg2_shape():
    """The network is E(3)-equivariant per molecule, so div_x b is unchanged by x -> R x + c (F = 128, L = 5,
    9 atoms: the BASELINE cfg-2 model; 64 conformers keep the 27 tangent passes short)."""
    from thermodynamic_interpolation_b200 import batch as B
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    torch.manual_seed(0)
    model = cPaiNN(n_features=128, score_layers=5, temp_length=100)
    perturb_(model, 1, 0.05)
    model = model.eval().to(DEV)
    batch = B.synthetic_ambient_batch(64, 9, seed=5).to(DEV)
    eng = model.engine()
    pb = eng.prepare(batch)
    x = batch.x0.clone()
    b0, d0 = eng.drift_div(pb, x, 0.3)
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    R = q.to(torch.float32).to(DEV)
    xr = x @ R.T + torch.tensor([0.3, -1.1, 0.7], device=DEV)
    b1, d1 = eng.drift_div(pb, xr, 0.3)
    assert torch.isfinite(d0).all() and float(d0.abs().max()) > 0
    _close(d1.cpu().numpy(), d0.cpu().numpy(), rtol=2e-3, atol_rel=2e-4, what="divergence under rotation+translation")
    _close(b1.cpu().numpy(), (b0 @ R.T).cpu().numpy(), rtol=1e-3, atol_rel=1e-4, what="drift equivariance")


def test_sampler_driver_writes_dlogps_file(tmp_path):
    """sample() with config.return_dlogp (mdqm9/sample_ambient.py:76-114): dlogps_<name>.npy holds the last-frame
    dlogp of every molecule, next to the samples the analysis scripts read (results_00031.py:173-201)."""
    import argparse
    from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    from thermodynamic_interpolation_b200.sample_ambient import sample
    torch.manual_seed(3)
    model = perturb_(cPaiNN(n_features=32, score_layers=2, temp_length=100), 4).eval()
    loader = [synthetic_ambient_batch(4, 9, seed=s) for s in (1, 2)]
    cfg = argparse.Namespace(seed=0, data_save_path=str(tmp_path), data_save_name="d", rtol=1e-4, atol=1e-4, n_steps=4,
                             return_dlogp=1)
    out = sample(cfg, model, loader, method="euler", verbose=False)
    dl = np.load(tmp_path / "dlogps_d.npy")
    assert dl.shape == (8,) and np.isfinite(dl).all() and np.abs(dl).max() > 0
    assert np.load(tmp_path / "samples_d.npy").shape == (8, 4, 9, 3)
    _, ref, _, _ = MoleculeIntegrator(model, method="euler", n_step=4, return_dlogp=True).rollout(
        synthetic_ambient_batch(4, 9, seed=2).to(DEV))
    np.testing.assert_allclose(dl[4:], ref[-1].cpu().numpy(), rtol=1e-6)
