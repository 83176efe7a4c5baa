"""Shared helpers for tests, oracle/make_golden.py, bench.py and smoke(): deterministic weights."""
from __future__ import annotations

import hashlib

import torch


from thermodynamic_interpolation_b200.synthetic import perturb_  # noqa: E402,F401  (re-exported for the tests)


def state_sha(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


# ---- golden fixtures -------------------------------------------------------------------------------
import os  # noqa: E402

import numpy as np  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def golden_batch(g):
    from thermodynamic_interpolation_b200.batch import MolBatch
    fields = {k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("in::")}
    if "x" not in fields:
        fields["x"] = fields["x0"].clone()
    return MolBatch(**fields)


def golden_model(g, device="cpu"):
    """Product-side parameter holder carrying the fixture's weights: stored tensors when present,
    otherwise the seeded recipe of oracle/make_golden.py pinned by the fixture's sha256."""
    kind = str(g["kind"])
    F, L = int(g["F"]), int(g["L"])
    if kind == "ambient":
        from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
        kw = dict(n_features=F, score_layers=L, temp_length=int(g["temp_length"]))
    else:
        from thermodynamic_interpolation_b200.latent.models.cpainn import cPaiNN
        kw = dict(n_features=F, score_layers=L, temp_length=int(g["temp_length"]),
                  temperatures=[int(t) for t in g["temperatures"]])
    seed = int(g["seed"])
    torch.manual_seed(seed)
    model = cPaiNN(**kw)
    stored = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w::")}
    if stored:
        model.load_state_dict(stored)
    else:
        perturb_(model, seed + 1, float(g["perturb"]))
    assert state_sha(model.state_dict()) == str(g["sha"]), "fixture weights could not be reproduced"
    return model.eval().to(device)


def oracle_hyper(g):
    from oracle import cpainn_oracle as co
    kind = str(g["kind"])
    kw = dict(n_features=int(g["F"]), score_layers=int(g["L"]), temp_length=int(g["temp_length"]), variant=kind)
    if kind == "latent":
        kw["temperatures"] = [int(t) for t in g["temperatures"]]
    return co.Hyper(**kw)


def oracle_temps(batch, hp):
    if hp.variant == "ambient":
        return dict(T0=batch.T0, T1=batch.T1)
    return dict(T=batch.T) if hp.n_temp_encoders == 1 else {}


def oracle_hp_sd(model):
    """Oracle-side (Hyper, state_dict) of a product model holder."""
    from oracle import cpainn_oracle as co
    hp_p = model.hyper
    hp = co.Hyper(n_features=hp_p.n_features, score_layers=hp_p.score_layers, temp_length=hp_p.temp_length,
                  time_length=hp_p.time_length, n_types=hp_p.n_types, temperatures=list(hp_p.temperatures),
                  variant=hp_p.variant)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    return hp, sd


def oracle_drift(model, batch, x, t):
    """Oracle drift for a product model holder + MolBatch (CPU tensors)."""
    from oracle import cpainn_oracle as co
    hp, sd = oracle_hp_sd(model)
    atoms = batch.atoms if hp.variant == "ambient" else batch.atom_number
    cpu = lambda v: v.detach().cpu()  # noqa: E731
    temps = {k: cpu(v) for k, v in oracle_temps(batch, hp).items()}
    return co.drift(sd, hp, cpu(x), t, cpu(atoms), cpu(batch.edge_index), cpu(batch.edge_type), **temps), hp, sd
