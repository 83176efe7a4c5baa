"""GPU tier: the CUDA path (libtib.so through the C ABI, driven by the drop-in classes) against
 (1) the frozen outputs of the unmodified reference (tests/golden/*.npz),
 (2) the CPU oracle (oracle/cpainn_oracle.py) on seeded inputs at sizes it finishes in seconds,
 (3) size-independent properties at BASELINE.json's full sizes.
Tolerances are fp32: BASELINE.json's bar is rtol 1e-4 on the state after 100 steps."""
import numpy as np
import pytest
import torch

from tests._util import golden_batch, golden_model, load_golden, oracle_drift, perturb_

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


def _integrator(kind):
    if kind == "ambient":
        from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    else:
        from thermodynamic_interpolation_b200.latent.integrators import MoleculeIntegrator
    return MoleculeIntegrator


def _wrapper(kind):
    if kind == "ambient":
        from thermodynamic_interpolation_b200.ambient.models.ode_wrapper import ODEWrapper
    else:
        from thermodynamic_interpolation_b200.latent.models.ode_wrapper import ODEWrapper
    return ODEWrapper


DRIFT_CASES = ["ambient_f32", "ambient_f128", "ambient_f256", "latent_multi_f64", "latent_single_f32",
               "latent_multi_f128", "latent_multi_f256"]


@pytest.mark.parametrize("name", DRIFT_CASES)
def test_drift_matches_reference_golden(name):
    """ODEWrapper.forward(t, x, batch) == the reference's on the same weights and inputs."""
    g = load_golden(name)
    kind = str(g["kind"])
    model = golden_model(g, DEV)
    batch = golden_batch(g).to(DEV)
    wrap = _wrapper(kind)(model)
    from thermodynamic_interpolation_b200 import _lib
    # n_features = 128 / 256 default to the tensor-core paths (split-f16 GEMMs, fp32-faithful: 2e-5 of max |drift|);
    # the fp32 CUDA-core kernels of the same width are held to the fp32 bound of the other fixtures
    modes = [(None, 5e-6)] if int(g["F"]) < 128 else [(None, 2e-5), (_lib.MATH_FP32_SIMT, 5e-6)]
    for mode, atol_rel in modes:
        if mode is not None:
            model.set_math(mode)
        for t, ref in zip(g["drift_t"], g["drift"]):
            args = (torch.tensor(float(t)), batch.x0.clone(), batch) + (([0],) if kind == "ambient" else ())
            out = wrap(*args)
            _close(out.cpu().numpy(), ref, rtol=1e-4, atol_rel=atol_rel, what=f"{name} drift t={t} math={mode}")


@pytest.mark.parametrize("name", DRIFT_CASES)
def test_forward_contract(name):
    """cPaiNN.forward(batch) writes batch.output from batch.x / batch.t (cpainn.py:93-115)."""
    g = load_golden(name)
    model = golden_model(g, DEV)
    batch = golden_batch(g).to(DEV)
    batch.x = batch.x0.clone()
    batch.t = torch.full((batch.x.shape[0],), float(g["drift_t"][1]), device=DEV)
    out = model(batch).output
    assert out.shape == batch.x.shape
    _close(out.cpu().numpy(), g["drift"][1], rtol=1e-4, atol_rel=5e-6, what=f"{name} forward")


# Whole-trajectory comparison only where the fixture is well conditioned: the coarse grids of the
# other fixtures (dt = 0.25 .. 0.33, random weights) amplify a 1e-7 perturbation of x0 to 1e-4 .. 4e-3
# within 3-4 steps (measured on the CPU oracle), so they are compared step by step below.
@pytest.mark.parametrize("name,method", [("ambient_f32", "euler"), ("ambient_f128", "euler"), ("latent_single_f32", "euler"),
                                         ("ambient_f32", "midpoint"), ("ambient_f32", "rk4")])
def test_fixed_grid_rollout_matches_reference_golden(name, method):
    g = load_golden(name)
    kind = str(g["kind"])
    ref = g[f"{method}_xts"]
    model = golden_model(g, DEV)
    batch = golden_batch(g).to(DEV)
    integ = _integrator(kind)(model, method=method, n_step=ref.shape[0])
    res = integ.rollout(batch)
    xts = res[0]
    assert xts.shape == ref.shape
    if kind == "ambient":
        assert len(res) == 4 and res[2] == (ref.shape[0] - 1) * {"euler": 1, "midpoint": 2, "rk4": 4}[method]
        assert torch.equal(res[3], batch.batch) and res[1].shape == (int(batch.ptr.numel() - 1),)
    else:
        assert len(res) == 3
    _close(xts.cpu().numpy(), ref, rtol=1e-4, atol_rel=2e-5, what=f"{name} {method} frames")


@pytest.mark.parametrize("name,method", [("ambient_f32", "euler"), ("ambient_f128", "euler"), ("ambient_f256", "euler"),
                                         ("latent_single_f32", "euler"), ("ambient_f32", "midpoint"),
                                         ("ambient_f32", "rk4"), ("ambient_f128", "midpoint"), ("ambient_f128", "rk4")])
def test_per_step_state_agreement_with_reference_golden(name, method):
    """Every single step of the reference trajectory is reproduced from the reference's own previous
    frame: rollout(start=t_k, end=t_k+1, n_step=2) from x0 = ref[k] must give ref[k+1]."""
    g = load_golden(name)
    kind = str(g["kind"])
    ref = g[f"{method}_xts"]
    model = golden_model(g, DEV)
    batch = golden_batch(g).to(DEV)
    times = torch.linspace(0.0, 1.0, ref.shape[0])
    for k in range(ref.shape[0] - 1):
        batch.x0 = torch.from_numpy(ref[k]).to(DEV)
        integ = _integrator(kind)(model, method=method, n_step=2, start=float(times[k]), end=float(times[k + 1]))
        xts = integ.rollout(batch)[0]
        assert torch.equal(xts[0], batch.x0)
        _close(xts[1].cpu().numpy(), ref[k + 1], rtol=1e-4, atol_rel=1e-5, what=f"{name} {method} step {k}")


@pytest.mark.parametrize("name", ["ambient_f32", "ambient_f128"])
def test_dopri5_rollout_matches_reference_golden(name):
    """Same step sequence as torchdiffeq on the reference drift: identical NFE, frames within fp32."""
    from thermodynamic_interpolation_b200 import _lib
    g = load_golden(name)
    model = golden_model(g, DEV).set_math(_lib.MATH_FP32_SIMT)    # the op-by-op fp32 path: same accept/reject sequence
    batch = golden_batch(g).to(DEV)
    integ = _integrator("ambient")(model, method="dopri5", n_step=6, atol=1e-5, rtol=1e-5)
    xts, dlogp, nfe, bvec = integ.rollout(batch)
    assert nfe == int(g["dopri5_nfe"]), (nfe, int(g["dopri5_nfe"]))
    _close(xts.cpu().numpy(), g["dopri5_xts"], rtol=1e-4, atol_rel=2e-5, what=f"{name} dopri5 frames")


def test_dopri5_tensor_core_mode_within_solver_tolerance():
    """With the tensor-core drift (error ~5e-6) an accept/reject decision may flip, so the bar is the solver's
    own: final state within a few x rtol of the reference trajectory, NFE within two attempts."""
    g = load_golden("ambient_f128")
    model = golden_model(g, DEV)                                  # library default for F = 128: split-f16 tcgen05
    integ = _integrator("ambient")(model, method="dopri5", n_step=6, atol=1e-5, rtol=1e-5)
    xts, dlogp, nfe, bvec = integ.rollout(golden_batch(g).to(DEV))
    model.engine().status()
    assert abs(nfe - int(g["dopri5_nfe"])) <= 12, (nfe, int(g["dopri5_nfe"]))
    ref = g["dopri5_xts"]
    err = np.abs(xts.cpu().numpy() - ref).max() / np.abs(ref).max()
    print(f"[parity] ambient_f128 dopri5 (tensor cores): nfe {nfe} vs {int(g['dopri5_nfe'])}, max rel-to-max err {err:.3e}")
    assert err < 2e-3


def test_hundred_step_state_agreement_vs_oracle():
    """BASELINE.json: per-step state agreement within rtol 1e-4 after 100 steps (fp32)."""
    from oracle import cpainn_oracle as co
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    torch.manual_seed(11)
    model = perturb_(cPaiNN(n_features=32, score_layers=3, temp_length=100), 12).eval()
    mb = synthetic_ambient_batch(6, [9, 9, 12, 25, 9, 17], seed=13)
    _, hp, sd = oracle_drift(model, mb, mb.x0, 0.0)
    # the reference rollout needs equal molecule sizes; the oracle's own Euler loop is ragged-safe
    ref, _, _ = co.rollout(sd, hp, mb.x0, mb.atoms, mb.edge_index, mb.edge_type, mb.ptr.tolist(),
                           method="euler", n_step=101, T0=mb.T0, T1=mb.T1)
    integ = _integrator("ambient")(model.to(DEV), method="euler", n_step=101)
    xts = integ.rollout(mb.to(DEV))[0]
    for k in (1, 10, 50, 100):
        _close(xts[k].cpu().numpy(), ref[k].numpy(), rtol=1e-4, atol_rel=2e-5, what=f"state after {k} steps")


@pytest.mark.parametrize("F,n_list", [(64, [2, 3, 9]), (128, [25, 9, 16]), (32, [9] * 37)])
def test_drift_vs_oracle_ragged(F, n_list):
    """Ragged and tiny molecules (2 atoms = one edge pair), tile tails, many molecules."""
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    torch.manual_seed(F)
    model = perturb_(cPaiNN(n_features=F, score_layers=2, temp_length=100), F + 1).eval()
    mb = synthetic_ambient_batch(len(n_list), n_list, seed=F + 2, T0=900.0, T1=400.0)
    ref, _, _ = oracle_drift(model, mb, mb.x0, 0.42)
    model = model.to(DEV)
    eng = model.engine()
    out = eng.drift(eng.prepare(mb.to(DEV)), mb.x0, 0.42)
    _close(out.cpu().numpy(), ref.numpy(), rtol=1e-4, atol_rel=5e-6, what=f"ragged F={F}")


def _cfg2(n_mol, seed=0, F=128, L=5):
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    torch.manual_seed(seed)
    model = perturb_(cPaiNN(n_features=F, score_layers=L, temp_length=100), seed + 1).eval().to(DEV)
    mb = synthetic_ambient_batch(n_mol, 9, seed=seed + 2).to(DEV)
    return model, mb


def test_full_size_block_diagonal_and_permutation():
    """BASELINE cfg 2 size (4096 x 9 atoms, F=128, L=5).  Properties the reference has:
    molecule i's drift is bit-identical whatever the other molecules are, and identical molecules give
    identical drifts wherever they sit in the batch."""
    model, mb = _cfg2(4096)
    eng = model.engine()
    pb = eng.prepare(mb)
    full = eng.drift(pb, mb.x0, 0.3).clone()
    assert torch.isfinite(full).all()
    # sub-batch of the first 64 molecules alone
    from thermodynamic_interpolation_b200.dist import shard_batch
    sub = shard_batch(mb, 0, 64)
    out_sub = eng.drift(eng.prepare(sub), sub.x0.contiguous(), 0.3)
    assert torch.equal(out_sub, full[: out_sub.shape[0]])
    # duplicate molecule 0 into slot 1000: same answer there
    x2 = mb.x0.clone()
    x2[9000:9009] = x2[0:9]
    out2 = eng.drift(pb, x2, 0.3)
    assert torch.equal(out2[9000:9009], out2[0:9])
    assert torch.equal(out2[0:9], full[0:9])


def test_full_size_rotation_equivariance_and_chirality():
    """b(x R^T) = b(x) R^T for proper rotations (to fp32), but NOT under reflection: the cross-product
    term makes the network chirality-sensitive (cpainn.py:296-300)."""
    model, mb = _cfg2(512)
    eng = model.engine()
    pb = eng.prepare(mb)
    x = mb.x0
    base = eng.drift(pb, x, 0.6).clone()
    gen = torch.Generator().manual_seed(5)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen, dtype=torch.float64))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    R = q.to(torch.float32).to(DEV)
    rot = eng.drift(pb, (x @ R.T).contiguous(), 0.6).clone()
    scale = base.abs().max()
    assert ((rot - base @ R.T).abs().max() / scale) < 2e-5
    M = torch.diag(torch.tensor([-1.0, 1.0, 1.0], device=DEV))
    mir = eng.drift(pb, (x @ M.T).contiguous(), 0.6)
    assert ((mir - base @ M.T).abs().max() / scale) > 1e-3


def test_euler_maruyama_extension():
    """EM mode has no reference oracle ("parity unpinned").  Required properties:
    eps = 0 (or no noise/score) is bit-identical to Euler; with injected noise and a second network as
    the score it matches the 5-line restatement built from this library's own drift calls."""
    model, mb = _cfg2(8, F=32, L=2)
    score, _ = _cfg2(8, seed=7, F=32, L=2)
    Integ = _integrator("ambient")
    T = 6
    base = Integ(model, method="euler", n_step=T).rollout(mb)[0]
    noise = torch.randn(T - 1, mb.x0.shape[0], 3, device=DEV, generator=torch.Generator(DEV).manual_seed(3))
    same = Integ(model, method="euler", n_step=T, eps=0.0, score=score).rollout(mb, noise=noise)[0]
    assert torch.equal(base, same)
    eps = 0.05
    em = Integ(model, method="euler", n_step=T, eps=eps, score=score).rollout(mb, noise=noise)[0]
    eng, seng = model.engine(), score.engine()
    pb = eng.prepare(mb)
    times = torch.linspace(0.0, 1.0, T)
    x = mb.x0.clone()
    for k in range(T - 1):
        dt = float(times[k + 1] - times[k])
        b = eng.drift(pb, x, float(times[k])).clone()
        s = seng.drift(pb, x, float(times[k])).clone()
        x = x + dt * b + (dt * eps) * s + float(np.sqrt(np.float32(2.0 * eps * dt))) * noise[k]
        assert torch.allclose(em[k + 1], x, rtol=1e-5, atol=1e-6)


def test_euler_maruyama_against_cpu_oracle_drifts():
    """The same property with nothing of this library on the reference side: x_{k+1} = x_k + dt (b + eps s) + sqrt(2 eps dt) z_k
     stepped on the CPU with the ORACLE's drift (oracle/cpainn_oracle.py, pinned against the
    unmodified reference) for both networks, on the noise injected into the CUDA rollout."""
    from oracle import cpainn_oracle as co
    from tests._util import oracle_hp_sd
    model, mb = _cfg2(6, F=32, L=2)
    score, _ = _cfg2(6, seed=7, F=32, L=2)
    T, eps = 9, 0.05
    noise = torch.randn(T - 1, mb.x0.shape[0], 3, generator=torch.Generator().manual_seed(11))
    em = _integrator("ambient")(model, method="euler", n_step=T, eps=eps, score=score).rollout(mb, noise=noise.to(DEV))[0].cpu()
    hp, sd = oracle_hp_sd(model)
    hps, sds = oracle_hp_sd(score)
    cpu = {k: mb[k].cpu() for k in ("x0", "atoms", "edge_index", "edge_type", "T0", "T1")}
    times = torch.linspace(0.0, 1.0, T)
    x = cpu["x0"].clone()
    worst = 0.0
    for k in range(T - 1):
        dt = float(times[k + 1] - times[k])
        kw = dict(T0=cpu["T0"], T1=cpu["T1"])
        b = co.drift(sd, hp, x, float(times[k]), cpu["atoms"], cpu["edge_index"], cpu["edge_type"], **kw)
        sc = co.drift(sds, hps, x, float(times[k]), cpu["atoms"], cpu["edge_index"], cpu["edge_type"], **kw)
        x = x + dt * b + (dt * eps) * sc + float(np.sqrt(np.float32(2.0 * eps * dt))) * noise[k]
        worst = max(worst, float((em[k + 1] - x).abs().max() / x.abs().max()))
    assert worst < 1e-4, worst


def test_step_euler_kernel_bit_exact():
    """K1: x + dt*b with product and sum rounded separately == torch eager (torchdiffeq's y0 + dt*f0)."""
    from thermodynamic_interpolation_b200.engine import DriftEngine
    gen = torch.Generator(DEV).manual_seed(1)
    for n in (3, 27, 4096 * 27 + 1):
        x = torch.randn(n, device=DEV, generator=gen)
        b = torch.randn(n, device=DEV, generator=gen)
        out = torch.empty_like(x)
        frame = torch.empty_like(x)
        DriftEngine.step_euler(_FakeEngine(), x, b, 0.005, out=out, frame=frame)
        ref = x + torch.tensor(0.005, device=DEV) * b
        assert torch.equal(out, ref) and torch.equal(frame, ref)


class _FakeEngine:
    """step_euler only needs the library, a device and a stream."""
    def __init__(self):
        from thermodynamic_interpolation_b200 import _lib
        self.lib = _lib.load()
        self.device = torch.device(DEV)

    def _stream(self):
        import ctypes as C
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)


def test_reweight_stats_match_reference_golden():
    from thermodynamic_interpolation_b200 import stats as S
    g = load_golden("stats")
    E0, E1, nd = (torch.from_numpy(g[k]).to(DEV) for k in ("E0", "E1", "neg_dlogp"))
    out = S.finalize(S.reweight_partials(E0, E1, nd).cpu())
    assert abs(out["ess"] - float(g["ess"])) < 1e-9 * float(g["ess"])
    assert abs(out["dF"] - float(g["dF"])) < 1e-10
    assert out["n"] == 4096


def test_adw_drift_and_divergence_match_reference_golden():
    from thermodynamic_interpolation_b200.adw.models.simple import FCNetMultiBeta
    g = load_golden("adw")
    torch.manual_seed(int(g["seed"]))
    model = perturb_(FCNetMultiBeta(1, 1, 256, 5).double(), int(g["seed"]) + 1, float(g["perturb"])).to(DEV)
    x0 = torch.from_numpy(g["in::x0"]).to(DEV)
    b0 = torch.from_numpy(g["in::beta0"]).to(DEV)
    b1 = torch.from_numpy(g["in::beta1"]).to(DEV)
    b, div = model.drift_div(x0, 0.3, b0, b1)
    np.testing.assert_allclose(b.cpu().numpy(), g["drift_t03"], rtol=1e-10, atol=1e-12)
    # the reference differentiates w.r.t. its fp32 state, so its divergence carries fp32 rounding
    np.testing.assert_allclose(div.cpu().numpy() * 1e-2, g["div_t03_scaled"], rtol=5e-7, atol=1e-12)


@pytest.mark.parametrize("method,n_step", [("euler", 11), ("dopri5", 9)])
def test_adw_rollout_matches_reference_golden(method, n_step):
    """StandardIntegrator.rollout (adw/thermo/integrators.py:33-68) on the reference's fp64 model + fp32 state."""
    from thermodynamic_interpolation_b200.adw.integrators import StandardIntegrator
    from thermodynamic_interpolation_b200.adw.models.simple import FCNetMultiBeta
    g = load_golden("adw")
    torch.manual_seed(int(g["seed"]))
    model = perturb_(FCNetMultiBeta(1, 1, 256, 5).double(), int(g["seed"]) + 1, float(g["perturb"])).to(DEV)
    x0 = torch.from_numpy(g["in::x0"]).to(DEV)
    b0 = torch.from_numpy(g["in::beta0"]).to(DEV)
    b1 = torch.from_numpy(g["in::beta1"]).to(DEV)
    integ = StandardIntegrator(model, method=method, n_step=n_step, atol=1e-4, rtol=1e-4, return_dlogp=True)
    x, dlogp = integ.rollout(x0, b0, b1)
    assert x.shape == (n_step, 64, 1) and dlogp.shape == (n_step, 64, 1)
    tol = 1e-5 if method == "euler" else 2e-4        # dopri5: the solver's own tolerance (rtol = atol = 1e-4)
    _close(x.cpu().numpy(), g[f"{method}_x"], rtol=tol, atol_rel=tol, what=f"adw {method} x")
    _close(dlogp.cpu().numpy(), g[f"{method}_dlogp"], rtol=tol * 10, atol_rel=tol, what=f"adw {method} dlogp")
    with pytest.raises(TypeError):
        StandardIntegrator(model, method="euler", n_step=3).rollout(x0, b0, b1)


def test_errors_are_loud():
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    model = cPaiNN(n_features=32, score_layers=1)
    mb = synthetic_ambient_batch(2, 9)
    mb.x = mb.x0.clone()
    mb.t = torch.zeros(18)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        model(mb)          # CPU model: no fallback
    model = model.to(DEV)
    bad = mb.clone().to(DEV)
    bad.edge_index = bad.edge_index[:, :-2]
    bad.edge_type = bad.edge_type[:-2]
    with pytest.raises(ValueError, match="complete digraph"):
        model(bad)


def test_sampler_driver_writes_reference_file_format(tmp_path):
    """sample() (driver loop of mdqm9/sample_ambient.py:18-119): files and shapes the reference analysis reads."""
    import argparse
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    from thermodynamic_interpolation_b200.sample_ambient import sample
    torch.manual_seed(3)
    model = perturb_(cPaiNN(n_features=32, score_layers=2, temp_length=100), 4).eval()
    loader = [synthetic_ambient_batch(5, 9, seed=s) for s in (1, 2)]
    cfg = argparse.Namespace(seed=0, data_save_path=str(tmp_path), data_save_name="t", rtol=1e-4, atol=1e-4, n_steps=7,
                             return_dlogp=0)
    out = sample(cfg, model, loader, method="euler", verbose=False)
    s = np.load(tmp_path / "samples_t.npy")
    assert s.shape == (10, 7, 9, 3) and np.array_equal(s, out["samples"])
    assert np.load(tmp_path / "latent_noises_t.npy").shape == (10, 9, 3)
    assert np.load(tmp_path / "latent_dlogps_t.npy").shape == (10,)
    x0 = torch.cat([b.x0.cpu() for b in loader]).reshape(10, 9, 3).numpy()
    np.testing.assert_array_equal(s[:, 0], x0)       # frame 0 is the start conformer, molecule by molecule


def test_latent_sampler_driver_writes_reference_file_format(tmp_path):
    """sample() of mdqm9/sample_latent.py:18-100: `samples_<name>_forward.npy` [n_mol, T, n, 3], frame 0 = x0 (the
    ambient dataset reads frames [:, 0] and [:, -1], mdqm9_ambient.py:173-199)."""
    import argparse
    from thermodynamic_interpolation_b200.batch import synthetic_latent_batch
    from thermodynamic_interpolation_b200.latent.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.sample_latent import sample
    torch.manual_seed(5)
    model = perturb_(cPaiNN(n_features=32, score_layers=2, temp_length=75), 6).eval()
    loader = [synthetic_latent_batch(4, 9, T=800, seed=s) for s in (1, 2, 3)]
    cfg = argparse.Namespace(seed=0, data_save_path=str(tmp_path), data_save_name="lat", rtol=1e-4, atol=1e-4, n_steps=5,
                             return_dlogp=0)
    out = sample(cfg, model, loader, method="euler", verbose=False, save_every=2)
    s = np.load(tmp_path / "samples_lat_forward.npy")
    assert s.shape == (12, 5, 9, 3) and np.array_equal(s, out["samples"]) and np.isfinite(s).all()
    x0 = torch.cat([b.x0.cpu() for b in loader]).reshape(12, 9, 3).numpy()
    np.testing.assert_array_equal(s[:, 0], x0)
    assert not (tmp_path / "dlogps_lat_forward.npy").exists()
