"""GPU: the training step of libtib.so (tib_train_loss_grad, tib_adam_step, tib_gemm_f16x3) against fixtures made by the
UNMODIFIED reference training step (oracle/make_golden.py::train_case) and against the CPU oracle (oracle/train_oracle.py)."""
import numpy as np
import pytest
import torch

from tests._util import golden_model, load_golden, oracle_hp_sd

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _lib():
    from thermodynamic_interpolation_b200 import _lib as L
    return L


# ---- the general split-f16 GEMM ------------------------------------------------------------------------------------------------
def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (200, 160, 96), (1000, 640, 256), (37, 32, 32)])
def test_gemm_forward_form_with_gather_and_bias(M, N, K):
    from thermodynamic_interpolation_b200.train import gemm_f16x3
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    src = torch.randn(50, K, generator=g).to(DEV)
    idx = torch.randint(0, 50, (M,), generator=g).to(DEV, torch.int32)
    W = (torch.randn(N, K, generator=g) * 0.3).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    out = gemm_f16x3(src, W, idx_a=idx, bias=bias)
    ref = src[idx.long()].double() @ W.double().T + bias.double()
    assert _rel(out, ref) < 2e-6
    # operand scales are exact (powers of two)
    out2 = gemm_f16x3(src, W, idx_a=idx, bias=bias, scale_a=0.0625, scale_b=4.0)
    assert _rel(out2, ref) < 2e-6


def test_gemm_data_gradient_form_accumulate_and_row_scatter():
    from thermodynamic_interpolation_b200.train import gemm_f16x3
    L = _lib()
    g = torch.Generator().manual_seed(1)
    R, O, I = 300, 640, 128
    dY = (torch.randn(R, O, generator=g) * 1e-6).to(DEV)            # gradient-sized values: needs the dynamic scale
    W = (torch.randn(O, I, generator=g) * 0.2).to(DEV)
    amax = dY.abs().max().reshape(1).clone()
    ref = dY.double() @ W.double()
    out = gemm_f16x3(dY, W, trans_b=True, amax_a=amax)
    assert _rel(out, ref) < 5e-6
    acc = torch.full((R, I), 1e-5, device=DEV)
    gemm_f16x3(dY, W, trans_b=True, amax_a=amax, out=acc, mode=L.GEMM_ACCUM)
    assert _rel(acc, ref + 1e-5) < 5e-6
    # scatter-add of result rows (d s[src] += ...)
    dst = torch.randint(0, 40, (R,), generator=g).to(DEV, torch.int32)
    sc = torch.zeros(40, I, device=DEV)
    gemm_f16x3(dY, W, trans_b=True, amax_a=amax, out=sc, c_idx=dst, mode=L.GEMM_ATOMIC)
    ref_sc = torch.zeros(40, I, dtype=torch.float64, device=DEV).index_add_(0, dst.long(), ref)
    assert _rel(sc, ref_sc) < 5e-6
    # without the dynamic scale the same product loses its low bits (f16 subnormals): the scale matters
    bad = gemm_f16x3(dY, W, trans_b=True)
    assert _rel(bad, ref) > 1e-4


@pytest.mark.parametrize("R", [64, 1000, 36864])
def test_gemm_weight_gradient_form_split_k(R):
    from thermodynamic_interpolation_b200.train import gemm_f16x3
    L = _lib()
    g = torch.Generator().manual_seed(R)
    O, I = 384, 128
    dY = (torch.randn(R, O, generator=g) * 1e-3).to(DEV)
    X = torch.randn(77, I, generator=g).to(DEV)
    idx = torch.randint(0, 77, (R,), generator=g).to(DEV, torch.int32)
    amax = dY.abs().max().reshape(1).clone()
    out = torch.zeros(O, 2 * I, device=DEV)                                   # a column block of a wider gradient matrix
    gemm_f16x3(dY, X, trans_a=True, trans_b=True, idx_b=idx, amax_a=amax, out=out[:, I:], mode=L.GEMM_ATOMIC, split_k=True,
               M=O, N=I, K=R)
    ref = dY.double().T @ X[idx.long()].double()
    assert _rel(out[:, I:], ref) < 3e-6
    assert float(out[:, :I].abs().max()) == 0.0


# ---- loss and gradients against the reference -----------------------------------------------------------------------------------
def _params_close_after_adam(name, p, ref, lr, n_steps):
    """Adam divides every gradient element by its own magnitude, so an element whose gradient is at rounding level (the split-K
    weight-gradient sums are atomic: their order differs from run to run) may legitimately move by up to lr per step in either
    direction.  Everything else must agree to 5e-6: at most 0.1 % of a tensor's entries may exceed that, none 2 lr per step."""
    diff = (p.detach().cpu() - ref).abs()
    assert float(diff.max()) < 2.0 * lr * n_steps + 1e-5, (name, float(diff.max()))
    assert float((diff > 5e-6).float().mean()) < 1e-3, (name, float((diff > 5e-6).float().mean()))


def _fixture_batches(g):
    from thermodynamic_interpolation_b200.batch import MolBatch
    out = []
    for i in (0, 1):
        out.append(MolBatch(**{k: torch.from_numpy(g[f"in{i}::{k}"]) for k in ("x", "T", "atoms", "edge_index", "edge_type", "batch", "ptr")}))
    return out


def _grad_dict(model, flat):
    from thermodynamic_interpolation_b200.engine import packed_keys
    out, off = {}, 0
    for k, shp in packed_keys(model.hyper):
        n = int(np.prod(shp))
        out[k] = flat[off:off + n].view(*shp).cpu()
        off += n
    assert off == flat.numel()
    return out


def _check_summary(name, tensor, summary, index, tol, scale=None):
    flat = tensor.detach().reshape(-1).to(torch.float64)
    ref_norm, ref_vals = float(summary[0]), summary[2:]
    scale = scale if scale is not None else max(float(np.abs(ref_vals).max()), ref_norm / np.sqrt(flat.numel()), 1e-30)
    assert abs(float(flat.norm()) - ref_norm) <= tol * max(ref_norm, 1e-30), (name, float(flat.norm()), ref_norm)
    err = float(np.abs(flat[torch.from_numpy(index)].numpy() - ref_vals).max())
    assert err <= tol * scale, (name, err, scale)


@pytest.mark.parametrize("name", ["train_f32", "train_f128_mixed", "train_f128"])
def test_loss_and_gradients_match_reference(name):
    from thermodynamic_interpolation_b200.train import TrainEngine, flatten, packed_parameters
    g = load_golden(name)
    model = golden_model(g, DEV)
    b0, b1 = _fixture_batches(g)
    eng = TrainEngine(model.hyper, DEV)
    tb = eng.prepare(b0, b1)
    w = flatten(packed_parameters(model))
    t, z = torch.from_numpy(g["t0"]), torch.from_numpy(g["z0"])
    loss, grad, b = eng.loss_and_grad(w, tb, t, z, gamma=str(g["gamma"]), want_b=True)
    eng.status()
    ref_loss = float(g["loss"][0])
    assert abs(float(loss) - ref_loss) < 5e-5 * max(1.0, abs(ref_loss)), (float(loss), ref_loss)
    grads = _grad_dict(model, grad)
    full = bool(int(g["full"]))
    worst = 0.0
    for k, gr in grads.items():
        if full:
            ref = torch.from_numpy(g["g::" + k])
            e = float((gr - ref).abs().max()) / max(float(ref.abs().max()), 1e-30)
            worst = max(worst, e)
            assert e < 3e-4, (k, e)
        else:
            _check_summary(k, gr, g["gs::" + k], g["gi::" + k], 3e-4)
    total = float(grad.double().norm())
    assert abs(total - float(g["grad_norm"][0])) < 2e-4 * float(g["grad_norm"][0])
    print(name, "loss", float(loss), "ref", ref_loss, "worst relative gradient error", worst)


def test_forward_of_both_passes_matches_oracle():
    from oracle import train_oracle as to
    from thermodynamic_interpolation_b200.train import TrainEngine, flatten, packed_parameters
    g = load_golden("train_f128")
    model = golden_model(g, DEV)
    hp, sd = oracle_hp_sd(model)
    b0, b1 = _fixture_batches(g)
    eng = TrainEngine(model.hyper, DEV)
    t, z = torch.from_numpy(g["t0"]), torch.from_numpy(g["z0"])
    _, _, b = eng.loss_and_grad(flatten(packed_parameters(model)), eng.prepare(b0, b1), t, z, gamma="sin2", want_b=True)
    with torch.no_grad():
        _, bp, bm = to.velocity_loss(sd, hp, b0.x, b1.x, t, z, b0.atoms, b0.edge_index, b0.edge_type, b0.T, b1.T, gamma="sin2", detach=True)
    ref = torch.cat([bp, bm])
    assert float((b.cpu() - ref).abs().max()) < 2e-5 * float(ref.abs().max())


@pytest.mark.parametrize("name", ["train_f32", "train_f128"])
def test_trainer_steps_match_reference_adam(name):
    """Two optimisation steps on the reference's draws: losses, clipped gradient norms and final parameters."""
    from thermodynamic_interpolation_b200.ambient.interpolants import LinearInterpolant
    from thermodynamic_interpolation_b200.train_ambient import Trainer
    g = load_golden(name)
    model = golden_model(g, DEV)
    b0, b1 = _fixture_batches(g)
    tr = Trainer(model, LinearInterpolant(a=1, gamma=str(g["gamma"])), lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
    for step in range(int(g["n_steps"])):
        loss = tr.step(b0, b1, t=torch.from_numpy(g[f"t{step}"]), z=torch.from_numpy(g[f"z{step}"]))
        assert abs(float(loss) - float(g["loss"][step])) < 1e-4 * max(1.0, abs(float(g["loss"][step]))), step
        assert abs(float(tr.last_grad_sqnorm.sqrt()) - float(g["grad_norm"][step])) < 5e-4 * float(g["grad_norm"][step])
    tr.sync_to_model()
    full = bool(int(g["full"]))
    for k, p in model.named_parameters():
        if p.dim() == 0:
            continue
        if full:
            _params_close_after_adam(k, p, torch.from_numpy(g["p::" + k]), float(g["lr"]), int(g["n_steps"]))
        else:
            flat = p.detach().cpu().reshape(-1).to(torch.float64)
            summary, index = g["ps::" + k], g["pi::" + k]
            assert abs(float(flat.norm()) - float(summary[0])) <= 2e-5 * max(float(summary[0]), 1e-30), k
            d = np.abs(flat[torch.from_numpy(index)].numpy() - summary[2:])
            # see _params_close_after_adam: an entry with a rounding-level gradient may move by lr per step
            assert d.max() < 2.0 * float(g["lr"]) * int(g["n_steps"]) + 1e-5 and (d > 2e-5).sum() <= 2, (k, d.max())


def test_reference_style_loop_with_torch_optimizer():
    """optim.zero_grad(); loss = loss_fn(batch0, batch1, model); loss.backward(); clip_grad_norm_; optim.step()
    (train_ambient.py:124-148) through the drop-in loss: same parameters as the reference after two steps."""
    from thermodynamic_interpolation_b200.ambient.interpolants import LinearInterpolant
    from thermodynamic_interpolation_b200.ambient.losses import StandardVelocityLoss
    g = load_golden("train_f32")
    model = golden_model(g, DEV).train()
    b0, b1 = _fixture_batches(g)
    loss_fn = StandardVelocityLoss(LinearInterpolant(a=1, gamma="sin2"), t_distr="uniform")
    optim = torch.optim.Adam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
    for step in range(int(g["n_steps"])):
        torch.manual_seed(1000 + int(g["seed"]) + step)          # the generator state the reference drew from
        optim.zero_grad()
        loss = loss_fn(b0, b1, model)
        loss.backward()
        assert abs(float(loss) - float(g["loss"][step])) < 1e-4
        norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        assert abs(float(norm) - float(g["grad_norm"][step])) < 5e-4 * float(g["grad_norm"][step])
        optim.step()
    for k, p in model.named_parameters():
        if p.dim() > 0:
            _params_close_after_adam(k, p, torch.from_numpy(g["p::" + k]), float(g["lr"]), int(g["n_steps"]))
    assert all(p.grad is None for p in model.parameters() if p.dim() == 0)       # the device_tracker dummies, as in the reference


def test_gradients_match_oracle_autograd_on_a_larger_batch():
    """32 molecules x 9 atoms, F = 128, L = 3, brownian gamma: CUDA adjoints vs torch autograd through the oracle."""
    from oracle import train_oracle as to
    from thermodynamic_interpolation_b200.batch import synthetic_train_batches
    from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model
    from thermodynamic_interpolation_b200.train import TrainEngine, flatten, packed_parameters
    model = seeded_ambient_model(128, 3, seed=21).to(DEV)
    hp, sd = oracle_hp_sd(model)
    b0, b1 = synthetic_train_batches(32, 9, 5)
    torch.manual_seed(3)
    t, z = to.draw_t_z([9] * 32)
    t = t.clamp(0.05, 0.95)                     # brownian gamma_dot is singular at the ends
    eng = TrainEngine(model.hyper, DEV)
    loss, grad, _ = eng.loss_and_grad(flatten(packed_parameters(model)), eng.prepare(b0, b1), t, z, gamma="brownian")
    eng.status()
    ref_loss, ref_grads, _, _ = to.loss_and_grads(sd, hp, b0.x, b1.x, t, z, b0.atoms, b0.edge_index, b0.edge_type, b0.T, b1.T, gamma="brownian")
    assert abs(float(loss) - float(ref_loss)) < 5e-5 * max(1.0, abs(float(ref_loss)))
    worst = 0.0
    for k, gr in _grad_dict(model, grad).items():
        ref = ref_grads[k]
        e = float((gr - ref).abs().max()) / max(float(ref.abs().max()), 1e-30)
        worst = max(worst, e)
        assert e < 3e-4, (k, e)
    print("worst relative gradient error vs oracle autograd", worst)


def test_f256_gradients_match_oracle_autograd():
    """The reference's second production width (10506_settings_no_900.json:14): 2 molecules x 25 atoms, F = 256, L = 2."""
    from oracle import train_oracle as to
    from thermodynamic_interpolation_b200.batch import synthetic_train_batches
    from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model
    from thermodynamic_interpolation_b200.train import TrainEngine, flatten, packed_parameters
    model = seeded_ambient_model(256, 2, seed=31).to(DEV)
    hp, sd = oracle_hp_sd(model)
    b0, b1 = synthetic_train_batches(2, 25, 7)
    torch.manual_seed(4)
    t, z = to.draw_t_z([25, 25])
    eng = TrainEngine(model.hyper, DEV)
    loss, grad, _ = eng.loss_and_grad(flatten(packed_parameters(model)), eng.prepare(b0, b1), t, z, gamma="sin2")
    eng.status()
    ref_loss, ref_grads, _, _ = to.loss_and_grads(sd, hp, b0.x, b1.x, t, z, b0.atoms, b0.edge_index, b0.edge_type, b0.T, b1.T, gamma="sin2")
    assert abs(float(loss) - float(ref_loss)) < 5e-5 * max(1.0, abs(float(ref_loss)))
    for k, gr in _grad_dict(model, grad).items():
        ref = ref_grads[k]
        assert float((gr - ref).abs().max()) <= 3e-4 * max(float(ref.abs().max()), 1e-30), k


def test_training_reduces_the_loss_and_rejects_bad_arguments():
    from thermodynamic_interpolation_b200.ambient.interpolants import LinearInterpolant
    from thermodynamic_interpolation_b200.batch import synthetic_train_batches
    from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model
    from thermodynamic_interpolation_b200.train_ambient import Trainer
    model = seeded_ambient_model(64, 2, seed=5).to(DEV)
    tr = Trainer(model, LinearInterpolant(a=1, gamma="sin2"), lr=1e-3)
    b0, b1 = synthetic_train_batches(24, 9, 11)
    g = torch.Generator().manual_seed(0)
    t = torch.rand(24, generator=g).repeat_interleave(9).reshape(-1, 1)
    z = torch.randn(24 * 9, 3, generator=g)
    losses = [float(tr.step(b0, b1, t=t, z=z)) for _ in range(30)]
    assert losses[-1] < losses[0] - 0.5 and all(np.isfinite(losses))
    before = [p.detach().clone() for p in model.parameters()]
    tr.sync_to_model()
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
    with pytest.raises(ValueError):
        tr.engine.loss_and_grad(tr.weights[:-1], tr.engine.prepare(b0, b1), t, z)
    with pytest.raises(NotImplementedError):
        Trainer(model, LinearInterpolant(a=1, gamma="sig_sum"))
    with pytest.raises(RuntimeError):
        from thermodynamic_interpolation_b200.train import TrainEngine
        TrainEngine(model.hyper, "cpu")


def test_cuda_graph_replay_equals_eager_steps():
    """Trainer(cuda_graph=True) replays the captured launches of both streams: same losses and weights as eager steps,
    also when the batch changes between replays."""
    from thermodynamic_interpolation_b200.ambient.interpolants import LinearInterpolant
    from thermodynamic_interpolation_b200.batch import synthetic_train_batches
    from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model
    from thermodynamic_interpolation_b200.train_ambient import Trainer
    ip = LinearInterpolant(a=1, gamma="sin2")
    eager = Trainer(seeded_ambient_model(128, 2, seed=5).to(DEV), ip)
    graph = Trainer(seeded_ambient_model(128, 2, seed=5).to(DEV), ip, cuda_graph=True)
    g = torch.Generator().manual_seed(0)
    for step in range(3):
        b0, b1 = synthetic_train_batches(8, 9, 40 + step)
        t = torch.rand(8, generator=g).repeat_interleave(9).reshape(-1, 1)
        z = torch.randn(72, 3, generator=g)
        le, lg = float(eager.step(b0, b1, t=t, z=z)), float(graph.step(b0, b1, t=t, z=z))
        assert abs(le - lg) < 1e-5 * max(1.0, abs(le)), (step, le, lg)
    assert float((eager.weights - graph.weights).abs().max()) < 3.5e-4      # Adam: rounding-level gradients move by <= lr per step
    assert float((eager.last_grad - graph.last_grad).abs().max()) < 1e-4 * float(eager.last_grad.abs().max())


def test_edge_cases_tiny_molecules_and_loud_errors():
    """One diatomic molecule (2 atoms, 2 edges, one pair) and a ragged pair of molecules against oracle autograd; wrong
    arguments fail loudly through the C ABI."""
    import ctypes as C
    from oracle import train_oracle as to
    from thermodynamic_interpolation_b200 import _lib
    from thermodynamic_interpolation_b200.batch import synthetic_train_batches
    from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model
    from thermodynamic_interpolation_b200.train import TrainEngine, flatten, packed_parameters
    model = seeded_ambient_model(64, 2, seed=41).to(DEV)
    hp, sd = oracle_hp_sd(model)
    eng = TrainEngine(model.hyper, DEV)
    w = flatten(packed_parameters(model))
    for sizes in ([2], [3, 2]):
        b0, b1 = synthetic_train_batches(len(sizes), sizes, 9)
        torch.manual_seed(6)
        t, z = to.draw_t_z(sizes)
        loss, grad, _ = eng.loss_and_grad(w, eng.prepare(b0, b1), t, z, gamma="sin2")
        eng.status()
        ref_loss, ref_grads, _, _ = to.loss_and_grads(sd, hp, b0.x, b1.x, t, z, b0.atoms, b0.edge_index, b0.edge_type, b0.T, b1.T, gamma="sin2")
        assert abs(float(loss) - float(ref_loss)) < 5e-5 * max(1.0, abs(float(ref_loss))), sizes
        for k, gr in _grad_dict(model, grad).items():
            ref = ref_grads[k]
            assert float((gr - ref).abs().max()) <= 3e-4 * max(float(ref.abs().max()), 1e-30) + 1e-9, (sizes, k)
    # loud errors: workspace too small, unknown gamma, latent descriptor
    lib = _lib.load()
    b0, b1 = synthetic_train_batches(2, 9, 1)
    tb = eng.prepare(b0, b1)
    t, z = to.draw_t_z([9, 9])
    t, z = t.to(DEV).reshape(-1).contiguous(), z.to(DEV)
    loss = torch.empty(1, dtype=torch.float64, device=DEV)
    grad = torch.empty_like(w)
    ws = torch.empty(4096, dtype=torch.uint8, device=DEV)
    pb = tb.pb
    cb = _lib.TrainBatch(n_mol=pb.n_mol, n_nodes=pb.n_nodes, n_edges=pb.n_edges, mol_ptr=pb.mol_ptr.data_ptr(), edge_ptr=pb.edge_ptr.data_ptr(),
                         atom_id=pb.atom_id.data_ptr(), edge_type=pb.edge_type.data_ptr(), temp0=pb.temp0.data_ptr(), temp1=pb.temp1.data_ptr(),
                         x0=tb.x0.data_ptr(), x1=tb.x1.data_ptr(), t=t.data_ptr(), z=z.data_ptr())

    def call(desc, ip, ws_bytes):
        return lib.tib_train_loss_grad(C.byref(desc), w.data_ptr(), C.byref(cb), C.byref(ip), loss.data_ptr(), grad.data_ptr(), None,
                                       ws.data_ptr(), ws_bytes, None)

    good_ip = _lib.Interpolant(gamma_kind=_lib.GAMMA_SIN2, a=1.0)
    assert call(eng.desc, good_ip, 4096) != 0 and b"workspace too small" in lib.tib_last_error()
    assert call(eng.desc, _lib.Interpolant(gamma_kind=7, a=1.0), 4096) != 0 and b"gamma" in lib.tib_last_error()
    from thermodynamic_interpolation_b200.engine import Hyper, model_desc
    latent = model_desc(Hyper(n_features=64, score_layers=2, variant="latent"))
    assert call(latent, good_ip, 4096) != 0 and b"ambient" in lib.tib_last_error()
    with pytest.raises(RuntimeError, match="ambient"):
        TrainEngine(Hyper(n_features=64, score_layers=2, variant="latent"), DEV)


@pytest.mark.parametrize("wd,max_norm", [(0.0, 1.0), (0.01, 0.0), (0.05, 0.5)])
def test_adam_step_matches_torch_optimizer(wd, max_norm):
    """tib_adam_step (clip_grad_norm_ + Adam with L2 weight decay, bias corrections) against torch.optim.Adam over 5 steps."""
    from thermodynamic_interpolation_b200.engine import Hyper
    from thermodynamic_interpolation_b200.train import TrainEngine
    eng = TrainEngine(Hyper(n_features=32, score_layers=1), DEV)
    g = torch.Generator().manual_seed(3)
    n = 10007
    w0 = torch.randn(n, generator=g)
    p = torch.nn.Parameter(w0.clone().to(DEV))
    opt = torch.optim.Adam([p], lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    w, m, v = w0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 6):
        grad = (torch.randn(n, generator=g) * 10.0 ** float(torch.randint(-3, 2, (1,), generator=g))).to(DEV)
        p.grad = grad.clone()
        if max_norm > 0:
            total = torch.nn.utils.clip_grad_norm_([p], max_norm)
        opt.step()
        sq = eng.adam_step(w, grad.clone(), m, v, step, lr=3e-3, weight_decay=wd, max_grad_norm=max_norm)
        if max_norm > 0:
            assert abs(float(sq.sqrt()) - float(total)) < 1e-5 * float(total)
        assert float((w - p.detach()).abs().max()) < 2e-6, step
    with pytest.raises(RuntimeError):
        eng.adam_step(w, grad, m, v, 0)
    with pytest.raises(ValueError):
        eng.adam_step(w, grad[:-1], m, v, 1)


def test_loss_without_grad_mode_returns_the_same_value():
    from thermodynamic_interpolation_b200.ambient.interpolants import LinearInterpolant
    from thermodynamic_interpolation_b200.ambient.losses import StandardVelocityLoss
    g = load_golden("train_f32")
    model = golden_model(g, DEV)
    b0, b1 = _fixture_batches(g)
    loss_fn = StandardVelocityLoss(LinearInterpolant(a=1, gamma="sin2"))
    t, z = torch.from_numpy(g["t0"]), torch.from_numpy(g["z0"])
    with torch.no_grad():
        l0 = loss_fn(b0, b1, model, t=t, z=z)
    l1 = loss_fn(b0, b1, model, t=t, z=z)
    assert not l0.requires_grad and l1.requires_grad
    assert abs(float(l0) - float(g["loss"][0])) < 1e-5 and abs(float(l0) - float(l1)) < 1e-6
    with pytest.raises(NotImplementedError):
        StandardVelocityLoss(LinearInterpolant(a=1, gamma="sig_sum"))(b0, b1, model)


def test_training_driver_writes_reference_checkpoint_files(tmp_path):
    """train_ambient.trainer: the reference's epoch loop and file names (mdqm9/train_ambient.py:96-176) around the native step;
    the written state_dicts load into a fresh model (same keys as the reference's checkpoints)."""
    import argparse
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.batch import synthetic_train_batches
    from thermodynamic_interpolation_b200.train_ambient import trainer
    cfg = argparse.Namespace(seed=0, n_features=32, score_layers=2, temp_length=100, a=1, gamma="sin2", t_distr="uniform",
                             learning_rate=2e-3, weight_decay=0, n_epochs=3, model_save_path=str(tmp_path), model_save_name="m0")
    data = [synthetic_train_batches(6, 9, seed=20 + i) for i in range(3)]

    def make_loaders(epoch):
        return [d[0] for d in data], [d[1] for d in data]

    hist = trainer(cfg, make_loaders, device=DEV, verbose=False)
    # (with fresh random times every batch the per-epoch loss of 18 molecules is too noisy to compare epochs;
    #  test_training_reduces_the_loss_and_rejects_bad_arguments checks the optimisation on fixed draws)
    assert len(hist["train_loss"]) == 3 and all(np.isfinite(hist["train_loss"])) and all(np.isfinite(hist["last_model_train_loss"]))
    assert all(b <= t + 1e-6 for b, t in zip(hist["epoch_best_loss"], hist["train_loss"]))
    assert hist["lr"][-1] == 2e-3
    for epoch in range(3):
        for tag in (f"m0_{epoch}_weights.pt", f"m0_best{epoch}_weights.pt"):
            sd = torch.load(str(tmp_path / "m0" / tag), map_location="cpu")
            fresh = cPaiNN(n_features=32, score_layers=2, temp_length=100)
            assert list(sd.keys()) == list(fresh.state_dict().keys())
            fresh.load_state_dict(sd)
    last = torch.load(str(tmp_path / "m0" / "m0_2_weights.pt"), map_location="cpu")
    best = torch.load(str(tmp_path / "m0" / "m0_best2_weights.pt"), map_location="cpu")
    assert any(not torch.equal(last[k], best[k]) for k in last)          # the best-of-epoch weights are an earlier snapshot
