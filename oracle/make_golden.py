"""TEST INFRASTRUCTURE ONLY (oracle/).  Generates tests/golden/*.npz by running the UNMODIFIED
reference modules from /root/reference (imported through oracle/stubs) on seeded synthetic inputs.
Runs in the build container only; the committed .npz files are what travels to the GPU box.

    python -m oracle.make_golden            # regenerate every fixture

Weights: small models are stored in the fixture; large ones are regenerated from
(torch.manual_seed(seed); construct; tests._util.perturb_(model, seed+1, 0.05)) and pinned by a
sha256 of the state_dict.  The product's parameter holders reproduce the reference's construction
order, so the same recipe yields the same weights without the reference (asserted here).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import ref_loader  # noqa: E402
from tests._util import perturb_, state_sha  # noqa: E402
from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch, synthetic_latent_batch, synthetic_train_batches  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")
PERTURB = 0.05


def to_ref_batch(tg, mb):
    b = tg.data.Batch()
    for k in mb.keys():
        b[k] = mb[k].clone() if torch.is_tensor(mb[k]) else mb[k]
    return b


def check_graph_contract(ns, n_atoms):
    """The reference's own AddRadiusGraph + AddBondGraph + Coalesce on one molecule must give the
    edge_index / edge_type the synthetic builder emits (mdqm9/data/mdqm9_ambient.py:160-170)."""
    tg = ns.torch_geometric
    mb = synthetic_ambient_batch(1, n_atoms, seed=3)
    n = n_atoms
    a = torch.arange(n - 1)
    bond_index = torch.stack([torch.cat([a, a + 1]), torch.cat([a + 1, a])])
    orders = torch.tensor([(1, 2, 1, 3)[i % 4] for i in range(n - 1)])
    d = tg.data.Data(x=mb.x0, x0=mb.x0, atoms=torch.arange(n), bond_index=bond_index, bonds=torch.cat([orders, orders]))
    b = tg.data.Batch.from_data_list([d])
    b = ns.utils.Coalesce()(ns.utils.AddBondGraph()(ns.utils.AddRadiusGraph(cutoff=1000)(b)))
    assert torch.equal(b.edge_index, mb.edge_index) and torch.equal(b.edge_type, mb.edge_type), "batch contract drifted"


def make_model(ref_cls, seed, **kw):
    torch.manual_seed(seed)
    m = ref_cls(**kw)
    perturb_(m, seed + 1, PERTURB)
    return m.eval()


def sd_arrays(sd):
    return {"w::" + k: v.detach().numpy() for k, v in sd.items()}


def ambient_case(ns, name, F, L, n_mol, n_atoms, seed, store_weights, n_frames=11, with_dopri=True, with_div=False):
    tg = ns.torch_geometric
    kw = dict(n_features=F, score_layers=L, temp_length=100)
    model = make_model(ns.ambient_cpainn.cPaiNN, seed, **kw)
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN as Mine
    mine = make_model(Mine, seed, **kw)
    assert state_sha(mine.state_dict()) == state_sha(model.state_dict()), "seeded weight recipe diverged"
    mb = synthetic_ambient_batch(n_mol, n_atoms, seed=seed + 10)
    out = dict(kind="ambient", F=F, L=L, temp_length=100, seed=seed, perturb=PERTURB, n_mol=n_mol,
               n_atoms=np.array([n_atoms] * n_mol if isinstance(n_atoms, int) else n_atoms), batch_seed=seed + 10,
               sha=state_sha(model.state_dict()))
    for k in ("x0", "T0", "T1", "atoms", "edge_index", "edge_type", "batch", "ptr"):
        out["in::" + k] = mb[k].numpy()
    wrap = ns.ambient_ode_wrapper.ODEWrapper(model)
    ts = [0.0, 0.37, 1.0]
    with torch.no_grad():
        drifts = [wrap(torch.tensor(t), mb.x0.clone(), to_ref_batch(tg, mb), [0]).numpy() for t in ts]
    out["drift_t"] = np.array(ts, dtype=np.float32)
    out["drift"] = np.stack(drifts)
    integ = ns.ambient_integrators.MoleculeIntegrator(model, method="euler", n_step=n_frames)
    xts, dlogp, nfe, bvec = integ.rollout(to_ref_batch(tg, mb))
    out["euler_xts"] = xts.numpy()
    out["euler_nfe"] = nfe
    if with_dopri:
        integ = ns.ambient_integrators.MoleculeIntegrator(model, method="dopri5", n_step=6, atol=1e-5, rtol=1e-5)
        xts, dlogp, nfe, _ = integ.rollout(to_ref_batch(tg, mb))
        out["dopri5_xts"] = xts.numpy()
        out["dopri5_nfe"] = nfe
        for meth in ("midpoint", "rk4"):
            integ = ns.ambient_integrators.MoleculeIntegrator(model, method=meth, n_step=5)
            out[f"{meth}_xts"] = integ.rollout(to_ref_batch(tg, mb))[0].numpy()
    if with_div:
        wrapd = ns.ambient_ode_wrapper.ODEWrapper(model, return_dlogp=True)
        b_, negdiv = wrapd(torch.tensor(0.37), (mb.x0.clone(), torch.zeros(n_mol)), to_ref_batch(tg, mb), [0])
        out["div_t037_scaled"] = (-negdiv).detach().numpy()     # divergence * 1e-2
        integ = ns.ambient_integrators.MoleculeIntegrator(model, method="euler", n_step=3, return_dlogp=True)
        xts, dlogp, nfe, _ = integ.rollout(to_ref_batch(tg, mb))
        out["euler_dlogp_xts"] = xts.detach().numpy()
        out["euler_dlogp"] = dlogp.detach().numpy()
    if store_weights:
        out.update(sd_arrays(model.state_dict()))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "drift |mean|", float(np.abs(out["drift"]).mean()), "bytes", os.path.getsize(os.path.join(GOLDEN, name + ".npz")))


def latent_case(ns, name, F, L, n_list, seed, temperatures, temp_length=75, store_weights=True):
    tg = ns.torch_geometric
    kw = dict(n_features=F, score_layers=L, temp_length=temp_length, temperatures=temperatures)
    model = make_model(ns.latent_cpainn.cPaiNN, seed, **kw)
    from thermodynamic_interpolation_b200.latent.models.cpainn import cPaiNN as Mine
    mine = make_model(Mine, seed, **kw)
    assert state_sha(mine.state_dict()) == state_sha(model.state_dict())
    multi = len(temperatures) > 1
    mb = synthetic_latent_batch(len(n_list), n_list, T=800 if multi else None, seed=seed + 10)
    out = dict(kind="latent", F=F, L=L, temp_length=temp_length, seed=seed, perturb=PERTURB,
               temperatures=np.array(temperatures), n_atoms=np.array(n_list), batch_seed=seed + 10,
               sha=state_sha(model.state_dict()))
    for k in mb.keys():
        out["in::" + k] = mb[k].numpy()
    wrap = ns.latent_ode_wrapper.ODEWrapper(model)
    ts = [0.0, 0.5, 1.0]
    with torch.no_grad():
        out["drift"] = np.stack([wrap(torch.tensor(t), mb.x0.clone(), to_ref_batch(tg, mb)).numpy() for t in ts])
    out["drift_t"] = np.array(ts, dtype=np.float32)
    if len(set(n_list)) == 1:   # the reference rollout needs equal sizes (latent/integrators.py:51)
        integ = ns.latent_integrators.MoleculeIntegrator(model, method="euler", n_step=9)
        xts, dlogp, bvec = integ.rollout(to_ref_batch(tg, mb))
        out["euler_xts"] = xts.numpy()
    if store_weights:
        out.update(sd_arrays(model.state_dict()))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "drift |mean|", float(np.abs(out["drift"]).mean()), "bytes", os.path.getsize(os.path.join(GOLDEN, name + ".npz")))


GRAD_SAMPLES = 48


def tensor_summary(t, gen):
    """(2-norm, sum, GRAD_SAMPLES entries at seeded positions) of one tensor - what the large training fixtures keep."""
    flat = t.detach().reshape(-1).to(torch.float64)
    idx = torch.randint(0, flat.numel(), (GRAD_SAMPLES,), generator=gen)
    return np.concatenate([[float(flat.norm()), float(flat.sum())], flat[idx].numpy()]), idx.numpy()


def train_case(ns, name, F, L, n_mol, n_atoms, seed, gamma="sin2", n_steps=2, lr=1e-4, weight_decay=0.0, full=False):
    """The reference's training step (mdqm9/train_ambient.py:124-148): StandardVelocityLoss(LinearInterpolant(a=1,
    gamma)) -> backward -> clip_grad_norm_(1) -> torch.optim.Adam, `n_steps` times on the same two batches.  Stored:
    the draws (t, z) of every step (re-drawn from the same generator state the loss consumed), the loss values, the
    un-clipped gradients of step 1 (full tensors, or per-tensor summaries) and the parameters after the last step."""
    tg = ns.torch_geometric
    kw = dict(n_features=F, score_layers=L, temp_length=100)
    model = make_model(ns.ambient_cpainn.cPaiNN, seed, **kw).train()
    sha0 = state_sha(model.state_dict())
    b0, b1 = synthetic_train_batches(n_mol, n_atoms, seed + 20)
    n_list = [n_atoms] * n_mol if isinstance(n_atoms, int) else list(n_atoms)
    interp = ns.ambient_interpolants.LinearInterpolant(a=1, gamma=gamma)
    loss_fn = ns.ambient_losses.StandardVelocityLoss(interpolant=interp, t_distr="uniform")
    optim = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)
    out = dict(kind="ambient", F=F, L=L, temp_length=100, seed=seed, perturb=PERTURB, n_mol=n_mol, gamma=gamma,
               n_atoms=np.array(n_list), batch_seed=seed + 20, sha=sha0, lr=lr, weight_decay=weight_decay, n_steps=n_steps,
               full=int(full))
    for k in ("x", "T", "atoms", "edge_index", "edge_type", "batch", "ptr"):
        out["in0::" + k] = b0[k].numpy()
        out["in1::" + k] = b1[k].numpy()
    from oracle import train_oracle
    gen = torch.Generator().manual_seed(seed + 99)
    names = [k for k, _ in model.named_parameters()]
    losses, norms = [], []
    for step in range(n_steps):
        torch.manual_seed(1000 + seed + step)
        state = torch.get_rng_state()
        t, z = train_oracle.draw_t_z(n_list)          # what the reference is about to draw
        torch.set_rng_state(state)
        out[f"t{step}"], out[f"z{step}"] = t.numpy(), z.numpy()
        optim.zero_grad()
        loss = loss_fn(to_ref_batch(tg, b0), to_ref_batch(tg, b1), model)
        loss.backward()
        losses.append(float(loss))
        if step == 0:
            for k, p in model.named_parameters():
                if p.grad is None:
                    continue
                if full:
                    out["g::" + k] = p.grad.detach().numpy().copy()
                else:
                    out["gs::" + k], out["gi::" + k] = tensor_summary(p.grad, gen)
        norms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), 1)))
        optim.step()
    out["loss"] = np.array(losses)
    out["grad_norm"] = np.array(norms)
    for k, p in model.named_parameters():
        if p.dim() == 0:
            continue
        if full:
            out["p::" + k] = p.detach().numpy().copy()
        else:
            out["ps::" + k], out["pi::" + k] = tensor_summary(p, gen)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "loss", losses, "grad norms", norms, "bytes", os.path.getsize(os.path.join(GOLDEN, name + ".npz")))


def adw_case(name="adw"):
    ns = ref_loader.load_adw()
    torch.manual_seed(5)
    model = ns.simple.FCNetMultiBeta(1, 1, 256, 5).double()
    perturb_(model, 6, 0.02)
    from thermodynamic_interpolation_b200.adw.models.simple import FCNetMultiBeta as Mine
    torch.manual_seed(5)
    mine = Mine(1, 1, 256, 5).double()
    perturb_(mine, 6, 0.02)
    assert state_sha(mine.state_dict()) == state_sha(model.state_dict())
    gen = torch.Generator().manual_seed(7)
    B = 64
    x0 = torch.randn(B, 1, generator=gen)
    beta0 = torch.full((B, 1), 1.0, dtype=torch.float64)
    beta1 = torch.full((B, 1), 1.25, dtype=torch.float64)
    out = dict(kind="adw", seed=5, perturb=0.02, sha=state_sha(model.state_dict()))
    out["in::x0"], out["in::beta0"], out["in::beta1"] = x0.numpy(), beta0.numpy(), beta1.numpy()
    wrap = ns.ode_wrapper.ODEWrapper(model, return_dlogp=True)
    b, negdiv = wrap(torch.tensor(0.3), (x0.clone(), torch.zeros(B, 1)), x0, beta0, beta1)
    out["drift_t03"] = b.detach().numpy()
    out["div_t03_scaled"] = (-negdiv).detach().numpy()
    for meth, n_step in (("euler", 11), ("dopri5", 9)):
        integ = ns.integrators.StandardIntegrator(model, method=meth, n_step=n_step, atol=1e-4, rtol=1e-4, return_dlogp=True)
        x, dlogp = integ.rollout(x0.clone(), beta0, beta1)
        out[f"{meth}_x"] = x.detach().numpy()
        out[f"{meth}_dlogp"] = dlogp.detach().numpy()
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "bytes", os.path.getsize(os.path.join(GOLDEN, name + ".npz")))


def stats_case(name="stats"):
    ns = ref_loader.load_analysis()
    rng = np.random.default_rng(11)
    n = 4096
    E0 = rng.normal(10.0, 2.0, n)
    E1 = E0 + rng.normal(0.3, 0.8, n)
    nd = rng.normal(0.0, 0.5, n)
    w = ns.ess.calc_ti_weights(E0, E1, nd)
    phis, keep = ns.free_energy.calc_phis_tfep(E0, E1, nd, k=None)
    out = dict(E0=E0, E1=E1, neg_dlogp=nd, weights=w, ess=ns.ess.calc_ESS(w),
               dF=ns.free_energy.calc_tfep_dF(phis, np.ones_like(phis)),
               keep_k3=ns.sensitivity.filter_iqr(w, k=3))
    # the bootstrap of results_00031.py:29-45 (that module loads trajectory files at import, so its loop is restated in
    # oracle/analysis_oracle.py around the reference's OWN calc_phis_tfep / calc_tfep_dF / filter_iqr)
    from oracle import analysis_oracle as ao
    for tag, k in (("none", None), ("k3", 3)):
        dF, ci = ao.bootstrap_dF(E0, E1, nd, ns.free_energy.calc_phis_tfep, ns.free_energy.calc_tfep_dF, n_bootstrap=200, k=k, seed=123)
        out[f"boot_dF_{tag}"], out[f"boot_ci_{tag}"] = dF, np.array(ci)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "ess", out["ess"], "dF", out["dF"])


def zmatrix_case(name="zmatrix"):
    """construct_z_matrix_batch of the unmodified reference on random conformers with a random valid
    reference-triplet table (every atom refers to three distinct earlier atoms where they exist)."""
    import importlib
    if ref_loader.REFERENCE_ROOT not in sys.path:
        sys.path.append(ref_loader.REFERENCE_ROOT)
    zm = importlib.import_module("mdqm9.analysis.utils.z_matrix")
    rng = np.random.default_rng(21)
    out = {}
    for tag, n_atoms, n_conf in (("a9", 9, 64), ("a25", 25, 16)):
        X = rng.normal(0.0, 1.0, (n_conf, n_atoms, 3)).astype(np.float32)
        order = rng.permutation(n_atoms).tolist()
        ref = []
        for a in range(n_atoms):
            placed = [order[i] for i in rng.permutation(a)]            # earlier atoms, random order
            filler = [o for o in order if o != order[a] and o not in placed]   # unused slots of rows 0..2
            ref.append([int(v) for v in (placed + filler)[:3]])
        z = zm.construct_z_matrix_batch(torch.tensor(X), ref, order).numpy()
        out[f"{tag}::X"], out[f"{tag}::order"], out[f"{tag}::ref"], out[f"{tag}::z"] = X, np.array(order), np.array(ref), z
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def main():
    """python -m oracle.make_golden [name ...]: regenerate every fixture, or only the named ones."""
    only = set(sys.argv[1:])
    want = lambda name: not only or name in only  # noqa: E731
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    ns = ref_loader.load_mdqm9()
    for n in (9, 25):
        check_graph_contract(ns, n)
    T8 = [300, 400, 500, 600, 700, 800, 900, 1000]
    if want("ambient_f32"):
        ambient_case(ns, "ambient_f32", F=32, L=2, n_mol=3, n_atoms=9, seed=0, store_weights=True, with_div=True)
    if want("ambient_f128"):
        ambient_case(ns, "ambient_f128", F=128, L=5, n_mol=4, n_atoms=9, seed=1, store_weights=False)
    if want("ambient_f256"):
        ambient_case(ns, "ambient_f256", F=256, L=2, n_mol=2, n_atoms=25, seed=2, store_weights=False, n_frames=4, with_dopri=False)
    if want("latent_multi_f64"):
        latent_case(ns, "latent_multi_f64", F=64, L=2, n_list=[9, 12, 25, 9], seed=3, temperatures=T8)
    if want("latent_single_f32"):
        latent_case(ns, "latent_single_f32", F=32, L=2, n_list=[9, 9, 9], seed=4, temperatures=[800])
    # round 2: the cfg-2 network over 100 consecutive Euler steps (north_star: "rtol 1e-4 after 100 steps"), the
    # reference's autograd divergence at F = 128 / L = 5, latent fixtures at the tensor-core widths
    if want("ambient_f128_100"):
        ambient_case(ns, "ambient_f128_100", F=128, L=5, n_mol=4, n_atoms=9, seed=1, store_weights=False, n_frames=101,
                     with_dopri=False)
    if want("ambient_f128_div"):
        ambient_case(ns, "ambient_f128_div", F=128, L=5, n_mol=3, n_atoms=9, seed=5, store_weights=False, n_frames=3,
                     with_dopri=False, with_div=True)
    if want("latent_multi_f128"):
        latent_case(ns, "latent_multi_f128", F=128, L=3, n_list=[9, 12, 9, 16], seed=6, temperatures=T8, store_weights=False)
    if want("latent_multi_f256"):
        latent_case(ns, "latent_multi_f256", F=256, L=2, n_list=[25, 9], seed=7, temperatures=T8, store_weights=False)
    # round 2: the training step: loss, gradients, clipping and Adam of the unmodified reference
    if want("train_f32"):
        train_case(ns, "train_f32", F=32, L=2, n_mol=3, n_atoms=9, seed=8, full=True)
    if want("train_f128"):
        train_case(ns, "train_f128", F=128, L=5, n_mol=12, n_atoms=9, seed=9)      # batch_size 12: 00031_settings_no_300.json:18
    if want("train_f128_mixed"):
        train_case(ns, "train_f128_mixed", F=128, L=2, n_mol=4, n_atoms=[9, 12, 7, 16], seed=10, gamma="brownian")
    if want("adw"):
        adw_case()
    if want("stats"):
        stats_case()
    if want("zmatrix"):
        zmatrix_case()


if __name__ == "__main__":
    main()
