"""CPU tier: host-side logic and the C-ABI surface (no GPU compute)."""
import ctypes as C
import re

import numpy as np
import pytest
import torch

from thermodynamic_interpolation_b200 import _lib, batch as B, dist as D
from thermodynamic_interpolation_b200.engine import DriftEngine, Hyper, PreparedBatch, pack_state_dict
from tests._util import golden_model, load_golden


def test_library_exports_every_declared_symbol():
    """Every function declared in include/tib.h is exported by libtib.so and bound in _lib.SYMBOLS."""
    import os
    hdr = open(os.path.join(os.path.dirname(_lib.HERE), "include", "tib.h")).read()
    declared = set(re.findall(r"\b(tib_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tib_model_desc", "tib_batch"}
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.tib_abi_version() == _lib.ABI_VERSION


@pytest.mark.parametrize("name", ["ambient_f32", "latent_multi_f64", "latent_single_f32"])
def test_state_dict_keys_match_reference_checkpoint_layout(name):
    """Key set AND key order equal the reference's state_dict (frozen in the fixture)."""
    g = load_golden(name)
    model = golden_model(g)
    ref_keys = [k[3:] for k in g if k.startswith("w::")]
    assert list(model.state_dict().keys()) == ref_keys


@pytest.mark.parametrize("variant,temps,F,L", [("ambient", None, 128, 5), ("latent", None, 64, 2), ("latent", [800], 32, 3)])
def test_packed_weight_count_matches_library(variant, temps, F, L):
    if variant == "ambient":
        from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
        m = cPaiNN(n_features=F, score_layers=L, temp_length=100)
    else:
        from thermodynamic_interpolation_b200.latent.models.cpainn import cPaiNN
        m = cPaiNN(n_features=F, score_layers=L, **({"temperatures": temps} if temps else {}))
    packed = pack_state_dict(m.state_dict(), m.hyper)
    hp = m.hyper
    desc = _lib.ModelDesc(abi_version=_lib.ABI_VERSION, variant=hp.c_variant, n_features=F, n_layers=L, n_types=25, n_edge_types=4)
    assert packed.size == _lib.load().tib_packed_weight_count(C.byref(desc))
    n_params = sum(p.numel() for p in m.parameters() if p.dim() > 0)
    assert packed.size == n_params


def test_reference_parameter_counts():
    """2,040,845 (ambient F=128 L=5), 2,024,458 (latent multi-T), 1,627,271 (single-T L=4)."""
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN as A
    from thermodynamic_interpolation_b200.latent.models.cpainn import cPaiNN as Lt
    count = lambda m: sum(p.numel() for p in m.parameters())  # noqa: E731
    assert count(A(n_features=128, score_layers=5)) == 2040845
    assert count(Lt(n_features=128, score_layers=5)) == 2024458
    assert count(Lt(n_features=128, score_layers=4, temperatures=[800])) == 1627271


def test_complete_digraph_is_coalesced_order():
    n_atoms = torch.tensor([3, 2, 4])
    ei, ptr, eptr = B.complete_digraph(n_atoms)
    assert ptr.tolist() == [0, 3, 5, 9] and eptr.tolist() == [0, 6, 8, 20]
    expect = [(i + o, j + o) for o, n in ((0, 3), (3, 2), (5, 4)) for i in range(n) for j in range(n) if i != j]
    assert ei.t().tolist() == [list(p) for p in expect]
    key = ei[0] * 100 + ei[1]
    assert bool((key[1:] > key[:-1]).all())


def test_synthetic_batch_contract():
    b = B.synthetic_ambient_batch(5, [9, 12, 9, 25, 10], seed=3)
    assert b.x0.shape == (65, 3) and b.x0.dtype == torch.float32
    for m in range(5):
        sl = slice(int(b.ptr[m]), int(b.ptr[m + 1]))
        assert torch.allclose(b.x0[sl].mean(0), torch.zeros(3), atol=1e-6)       # per-molecule centred
        assert b.atoms[sl].tolist() == list(range(sl.stop - sl.start))           # atoms = arange(n)
    assert b.edge_index.shape[1] == sum(n * (n - 1) for n in [9, 12, 9, 25, 10])
    assert set(b.edge_type.unique().tolist()) <= {0, 1, 2, 3}
    # bonds are symmetric: type(i->j) == type(j->i)
    n = 9
    et = b.edge_type[: n * (n - 1)]
    dense = torch.zeros(n, n, dtype=torch.long)
    dense[~torch.eye(n, dtype=torch.bool)] = et
    assert torch.equal(dense, dense.t())
    lb = B.synthetic_latent_batch(3, 9, T=800, seed=1)
    assert lb.T.dtype == torch.long and bool((lb.x == 0).all()) and lb.atom_number.dtype == torch.long


def test_prepared_batch_validation():
    hp = Hyper(n_features=32, score_layers=1)
    b = B.synthetic_ambient_batch(3, [9, 5, 7], seed=0)
    pb = PreparedBatch(b, hp, torch.device("cpu"))
    assert (pb.n_mol, pb.n_nodes, pb.n_edges, pb.max_atoms) == (3, 21, 72 + 20 + 42, 9)
    assert pb.mol_ptr.dtype == torch.int32 and pb.edge_ptr.dtype == torch.int64 and pb.edge_type.dtype == torch.uint8
    bad = b.clone()
    bad.edge_index = bad.edge_index[:, :-1]
    bad.edge_type = bad.edge_type[:-1]
    with pytest.raises(ValueError, match="complete digraph"):
        PreparedBatch(bad, hp, torch.device("cpu"))
    bad = b.clone()
    bad.edge_index = bad.edge_index.flip(0)          # (dst,src) order is not the coalesced order
    with pytest.raises(ValueError, match="complete digraph"):
        PreparedBatch(bad, hp, torch.device("cpu"))
    bad = b.clone()
    bad.edge_type = bad.edge_type + 4
    with pytest.raises(ValueError, match="edge_type"):
        PreparedBatch(bad, hp, torch.device("cpu"))
    bad = b.clone()
    bad.atoms = bad.atoms + 30
    with pytest.raises(ValueError, match="atom ids"):
        PreparedBatch(bad, hp, torch.device("cpu"))
    empty = B.MolBatch(atoms=torch.zeros(0, dtype=torch.long), batch=torch.zeros(0, dtype=torch.long))
    with pytest.raises(ValueError, match="empty"):
        PreparedBatch(empty, hp, torch.device("cpu"))


def test_no_cpu_fallback():
    """The product path refuses to run without a CUDA device instead of silently computing on the host."""
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    m = cPaiNN(n_features=32, score_layers=1)
    b = B.synthetic_ambient_batch(2, 9)
    b.t = torch.zeros(18)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(b)
    with pytest.raises(RuntimeError, match="CUDA"):
        MoleculeIntegrator(m, method="euler", n_step=3).rollout(b)
    with pytest.raises(RuntimeError, match="CUDA"):
        MoleculeIntegrator(m, method="euler", n_step=3, return_dlogp=True).rollout(b)


def test_library_argument_errors_are_reported():
    lib = _lib.load()
    desc = _lib.ModelDesc(abi_version=_lib.ABI_VERSION, variant=0, n_features=48, n_layers=1, n_types=25, n_edge_types=4)
    h = C.c_void_p()
    w = np.zeros(8, dtype=np.float32)
    assert lib.tib_model_create(C.byref(h), C.byref(desc), w.ctypes.data_as(C.c_void_p), 8, 0) != 0
    assert b"n_features" in lib.tib_last_error()
    desc.n_features = 32
    assert lib.tib_model_create(C.byref(h), C.byref(desc), w.ctypes.data_as(C.c_void_p), 8, 0) != 0
    assert b"mismatch" in lib.tib_last_error()
    desc.abi_version = 99
    assert lib.tib_model_create(C.byref(h), C.byref(desc), w.ctypes.data_as(C.c_void_p), 8, 0) != 0
    assert b"ABI" in lib.tib_last_error()


def test_shard_range_and_shard_batch():
    for n, w in ((10, 3), (7, 8), (4096, 8), (1, 2)):
        spans = [D.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    b = B.synthetic_ambient_batch(5, [9, 12, 9, 25, 10], seed=3)
    parts = [D.shard_batch(b, r, 2) for r in range(2)]
    assert sum(p.x0.shape[0] for p in parts) == b.x0.shape[0]
    assert torch.equal(torch.cat([p.x0 for p in parts]), b.x0)
    assert torch.equal(torch.cat([p.edge_type for p in parts]), b.edge_type)
    hp = Hyper(n_features=32, score_layers=1)
    for p in parts:                                 # each shard is a valid self-contained batch
        pb = PreparedBatch(p, hp, torch.device("cpu"))
        assert pb.n_mol == p.num_graphs


def test_regroup_frames_matches_reference_loop():
    """sample_ambient.py:93 regroups frames with a Python loop over molecules; ours is one reshape."""
    from thermodynamic_interpolation_b200.sample_ambient import regroup_frames
    rng = np.random.default_rng(0)
    T, B, n = 5, 7, 9
    xts = rng.normal(size=(T, B * n, 3)).astype(np.float32)
    batch_idx = np.repeat(np.arange(B), n)
    ref = np.array([xts[:, batch_idx == i] for i in range(batch_idx.max() + 1)])
    got = regroup_frames(xts, batch_idx)
    assert got.shape == (B, T, n, 3) and np.array_equal(got, ref)
    with pytest.raises(ValueError):
        regroup_frames(xts[:, :-1], batch_idx[:-1])


# ---- training host side (no GPU compute) -------------------------------------------------------------------------------------
def test_packed_parameters_follow_the_library_packing_order():
    """The flat weight / gradient vector of tib_train_loss_grad uses the order of tib_packed_weight_count: the parameter list,
    its total length and the reference's state_dict shapes must agree."""
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.engine import model_desc, packed_keys
    from thermodynamic_interpolation_b200.train import flatten, packed_parameters
    model = cPaiNN(n_features=64, score_layers=3, temp_length=100)
    params = packed_parameters(model)
    keys = packed_keys(model.hyper)
    sd = model.state_dict()
    assert len(params) == len(keys)
    for p, (k, shape) in zip(params, keys):
        assert tuple(p.shape) == tuple(shape) == tuple(sd[k].shape) and p.data_ptr() == sd[k].data_ptr(), k
    desc = model_desc(model.hyper)
    assert flatten(params).numel() == _lib.load().tib_packed_weight_count(C.byref(desc))
    assert np.array_equal(flatten(params).numpy(), pack_state_dict(sd, model.hyper))
    # the scalar device_tracker dummies of the reference are not part of the packed vector (they get no gradient)
    assert all(p.dim() > 0 for p in params) and any(p.dim() == 0 for p in model.parameters())


def test_training_is_cuda_only_and_draws_follow_the_reference_order():
    from thermodynamic_interpolation_b200.ambient import interpolants, losses
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.train import TrainEngine
    from oracle import train_oracle as to
    model = cPaiNN(n_features=32, score_layers=1, temp_length=100)
    with pytest.raises(RuntimeError, match="CUDA"):
        TrainEngine(model.hyper, "cpu")
    torch.manual_seed(7)
    t = losses.draw_times([9, 4, 9])
    torch.manual_seed(7)
    t_ref, _ = to.draw_t_z([9, 4, 9])
    assert torch.equal(t, t_ref) and t.shape == (22, 1)
    with pytest.raises(ValueError):
        losses.draw_times([3], "gaussian")
    # the drop-in interpolant's callables equal the oracle's restatement of interpolants.py:71-82
    tt = torch.linspace(0.05, 0.95, 7).unsqueeze(1)
    for kind in ("sin2", "brownian"):
        ip = interpolants.LinearInterpolant(a=1, gamma=kind)
        g, gd = to.gamma_fns(kind, 1.0)
        assert torch.equal(ip.gamma(tt), g(tt)) and torch.equal(ip.gamma_dot(tt), gd(tt)) and ip.kind == kind
    x0, x1 = torch.randn(7, 3), torch.randn(7, 3)
    ip = interpolants.LinearInterpolant(a=1, gamma="sin2")
    assert torch.allclose(ip.It(tt, x0, x1), (1 - tt) * x0 + tt * x1) and torch.equal(ip.dtIt(tt, x0, x1), -1.0 * x0 + 1.0 * x1)
    with pytest.raises(NotImplementedError):
        interpolants.LinearInterpolant(gamma="cubic")
