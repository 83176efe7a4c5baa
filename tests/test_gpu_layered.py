"""GPU tier: the layered tensor-core path (csrc/tc_chain.cuh + csrc/layered.cuh) - the F = 256 drift and the exact
divergence with its tangent GEMMs on tcgen05.  Checked against the reference's frozen outputs (tests/golden), the CPU
oracle (autograd divergence) and this library's fp32 CUDA-core kernels."""
import numpy as np
import pytest
import torch

from tests._util import golden_batch, golden_model, load_golden, oracle_drift, oracle_hp_sd, oracle_temps, perturb_
from thermodynamic_interpolation_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def _model(F, L, seed, variant="ambient"):
    torch.manual_seed(seed)
    if variant == "ambient":
        from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
        return perturb_(cPaiNN(n_features=F, score_layers=L, temp_length=100), seed + 1).eval()
    from thermodynamic_interpolation_b200.latent.models.cpainn import cPaiNN
    return perturb_(cPaiNN(n_features=F, score_layers=L, temp_length=75), seed + 1).eval()


@pytest.mark.parametrize("F,n_list", [(128, [9] * 40), (128, [25, 9, 16, 2, 3, 12]), (256, [9] * 17), (256, [25, 9, 16, 2, 3, 12])])
def test_layered_drift_vs_fp32_and_oracle(F, n_list):
    """MLP chains on tcgen05 + fp32 scatter: same drift as the fp32 CUDA-core path and as the CPU oracle."""
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    model = _model(F, 3, 31)
    mb = synthetic_ambient_batch(len(n_list), n_list, seed=33, T0=900.0, T1=400.0)
    ref, _, _ = oracle_drift(model, mb, mb.x0, 0.42)
    model = model.to(DEV)
    eng = model.engine()
    pb = eng.prepare(mb.to(DEV))
    x = mb.x0.to(DEV)
    model.set_math(_lib.MATH_FP32_SIMT)
    simt = eng.drift(pb, x, 0.42).cpu()
    model.set_math(_lib.MATH_F16X3_LAYERED if F == 128 else _lib.MATH_F16X3_TC)
    out = eng.drift(pb, x, 0.42)
    eng.status()
    out2 = eng.drift(pb, x, 0.42)
    assert torch.equal(out, out2), "repeated calls must be bit-identical"
    e1, e2 = _rel(out.cpu().numpy(), simt.numpy()), _rel(out.cpu().numpy(), ref.numpy())
    print(f"[layered] F={F} {len(n_list)} molecules: vs fp32 kernels {e1:.3e}, vs oracle {e2:.3e}")
    assert e1 < 4e-5 and e2 < 4e-5      # small ragged batches: worst case of the fp32-vs-split-f16 fuzz is 2.5e-5


def test_layered_f256_matches_reference_golden():
    """F = 256 (the reference's 10506 system, config/ambient/10506_settings_no_900.json:14) on tensor cores against
    the unmodified reference's frozen drift and Euler frames."""
    from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    from thermodynamic_interpolation_b200.ambient.models.ode_wrapper import ODEWrapper
    g = load_golden("ambient_f256")
    model = golden_model(g, DEV)
    assert model.engine() is not None
    model.set_math(_lib.MATH_F16X3_TC)
    batch = golden_batch(g).to(DEV)
    wrap = ODEWrapper(model)
    for t, ref in zip(g["drift_t"], g["drift"]):
        out = wrap(torch.tensor(float(t)), batch.x0.clone(), batch, [0])
        model.engine().status()
        err = _rel(out.cpu().numpy(), ref)
        print(f"[layered] ambient_f256 drift t={t}: {err:.3e}")
        assert err < 2e-5
    # per step from the reference's own previous frame (the coarse 4-frame grid of this fixture amplifies 1e-7 to 1e-3
    # over the whole trajectory - see tests/test_gpu_parity.py)
    ref = g["euler_xts"]
    times = torch.linspace(0.0, 1.0, ref.shape[0])
    for k in range(ref.shape[0] - 1):
        batch.x0 = torch.from_numpy(ref[k]).to(DEV)
        integ = MoleculeIntegrator(model, method="euler", n_step=2, start=float(times[k]), end=float(times[k + 1]))
        xts = integ.rollout(batch)[0]
        err = _rel(xts[1].cpu().numpy(), ref[k + 1])
        print(f"[layered] ambient_f256 euler step {k}: {err:.3e}")
        assert err < 3e-5      # per-step bound of north_star: 1e-4


def _oracle_div(model, batch, x, t):
    from oracle import cpainn_oracle as co
    hp, sd = oracle_hp_sd(model)
    atoms = batch.atoms if hp.variant == "ambient" else batch.atom_number
    cpu = lambda v: v.detach().cpu()  # noqa: E731
    temps = {k: cpu(v) for k, v in oracle_temps(batch, hp).items()}
    return co.divergence(sd, hp, cpu(x), t, cpu(atoms), cpu(batch.edge_index), cpu(batch.edge_type),
                         batch.ptr.tolist(), **temps)


@pytest.mark.parametrize("F,L,n_list", [(128, 2, [9] * 6), (128, 3, [5, 9, 3, 12, 2]), (256, 2, [9, 7, 4])])
def test_tc_divergence_vs_fp32_tangents_and_oracle(F, L, n_list):
    """Tangent GEMMs on tcgen05 (one tangent for the w MLP, none for the first phi MLP) against the dual-number fp32
    kernels of this library and the oracle's autograd divergence (ode_wrapper.py:59-91)."""
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    model = _model(F, L, 41)
    mb = synthetic_ambient_batch(len(n_list), n_list, seed=43, T0=800.0, T1=300.0)
    div_ref = _oracle_div(model, mb, mb.x0, 0.37)
    model = model.to(DEV)
    eng = model.engine()
    pb = eng.prepare(mb.to(DEV))
    x = mb.x0.to(DEV)
    model.set_math(_lib.MATH_FP32_SIMT)
    b0, d0 = eng.drift_div(pb, x, 0.37)
    model.set_math(_lib.MATH_F16X3_TC)
    b1, d1 = eng.drift_div(pb, x, 0.37)
    eng.status()
    eb, ed, eo = _rel(b1.cpu().numpy(), b0.cpu().numpy()), _rel(d1.cpu().numpy(), d0.cpu().numpy()), _rel(d1.cpu().numpy(), div_ref.numpy())
    print(f"[layered] F={F} L={L}: drift vs fp32 {eb:.3e}, divergence vs fp32 tangents {ed:.3e}, vs oracle autograd {eo:.3e}")
    assert eb < 2e-5 and ed < 1e-4 and eo < 1e-4


def test_tc_divergence_matches_reference_golden():
    """The reference's own divergence (3n autograd passes), frozen in tests/golden/ambient_f32.npz is F = 32 (fp32 kernels);
    the F = 128 fixture made from the unmodified reference pins the tensor-core tangents."""
    from thermodynamic_interpolation_b200.ambient.models.ode_wrapper import ODEWrapper
    g = load_golden("ambient_f128_div")
    model = golden_model(g, DEV)
    model.set_math(_lib.MATH_F16X3_TC)
    batch = golden_batch(g).to(DEV)
    x0 = batch.x0.clone()
    b, negdiv = ODEWrapper(model, return_dlogp=True)(torch.tensor(0.37), (x0, torch.zeros(int(g["n_mol"]), device=DEV)), batch, [0])
    model.engine().status()
    eb = _rel(b.cpu().numpy(), g["drift"][1])      # drift_t = [0, 0.37, 1]
    ed = _rel((-negdiv).cpu().numpy(), g["div_t037_scaled"])
    print(f"[layered] ambient_f128_div vs reference: drift {eb:.3e}, divergence {ed:.3e}")
    assert eb < 2e-5 and ed < 1e-5
