"""CPU: pins oracle/train_oracle.py (loss, autograd gradients, clipping, Adam) against fixtures made by the UNMODIFIED
reference training step (oracle/make_golden.py::train_case; mdqm9/train_ambient.py:124-148)."""
import numpy as np
import pytest
import torch

from oracle import train_oracle as to
from tests._util import golden_model, load_golden, oracle_hp_sd


def fixture_batches(g):
    f = lambda i, k: torch.from_numpy(g[f"in{i}::{k}"])  # noqa: E731
    return {k: f(0, k) for k in ("x", "T", "atoms", "edge_index", "edge_type", "ptr")}, {k: f(1, k) for k in ("x", "T")}


def check_against_summary(name, tensor, summary, index, tol, scale=None):
    flat = tensor.detach().reshape(-1).to(torch.float64)
    ref_norm, ref_vals = float(summary[0]), summary[2:]
    scale = scale if scale is not None else max(float(np.abs(ref_vals).max()), ref_norm / np.sqrt(flat.numel()), 1e-30)
    assert abs(float(flat.norm()) - ref_norm) <= tol * max(ref_norm, 1e-30), name
    assert float(np.abs(flat[torch.from_numpy(index)].numpy() - ref_vals).max()) <= tol * scale, name


@pytest.mark.parametrize("name", ["train_f32", "train_f128_mixed", "train_f128"])
def test_oracle_training_steps_match_reference(name):
    g = load_golden(name)
    model = golden_model(g)
    hp, sd = oracle_hp_sd(model)
    b0, b1 = fixture_batches(g)
    params = {k: v.clone() for k, v in sd.items() if v.is_floating_point() and v.dim() > 0}
    opt = to.Adam(lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
    full = bool(int(g["full"]))
    for step in range(int(g["n_steps"])):
        t, z = torch.from_numpy(g[f"t{step}"]), torch.from_numpy(g[f"z{step}"])
        loss, grads, _, _ = to.loss_and_grads({**sd, **params}, hp, b0["x"], b1["x"], t, z, b0["atoms"], b0["edge_index"],
                                              b0["edge_type"], b0["T"], b1["T"], gamma=str(g["gamma"]))
        assert abs(float(loss) - float(g["loss"][step])) < 2e-5 * max(1.0, abs(float(g["loss"][step])))
        grads = {k: v for k, v in grads.items() if v is not None and k in params}
        if step == 0:
            for k, gr in grads.items():
                if full:
                    ref = torch.from_numpy(g["g::" + k])
                    assert float((gr - ref).abs().max()) <= 2e-4 * max(float(ref.abs().max()), 1e-30), k
                else:
                    check_against_summary(k, gr, g["gs::" + k], g["gi::" + k], 2e-4)
        total = to.clip_grad_norm(grads, 1.0)
        assert abs(total - float(g["grad_norm"][step])) < 2e-4 * float(g["grad_norm"][step])
        opt.step(params, grads)
    for k, p in params.items():
        if full:
            assert float((p - torch.from_numpy(g["p::" + k])).abs().max()) < 2e-6, k
        else:
            check_against_summary(k, p, g["ps::" + k], g["pi::" + k], 1e-5, scale=1.0)


def test_draws_follow_the_reference_order():
    """One rand(1) per molecule repeated over its atoms, then randn(N, 3) (losses.py:46-47, interpolants.py:29)."""
    torch.manual_seed(5)
    t, z = to.draw_t_z([3, 2])
    torch.manual_seed(5)
    a, b = torch.rand(1), torch.rand(1)
    zz = torch.randn(5, 3)
    assert torch.equal(t.squeeze(1), torch.cat([a.repeat(3), b.repeat(2)])) and torch.equal(z, zz)
