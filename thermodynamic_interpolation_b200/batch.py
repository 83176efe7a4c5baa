"""Batch container + synthetic MDQM9-shaped batches.

The reference hands a `torch_geometric.data.Batch` to `MoleculeIntegrator.rollout`
(mdqm9/sample_ambient.py:73-88).  The hot path only reads attributes from it, so any object with
the same attribute names works here (a real PyG Batch included); `MolBatch` is the dependency-free
stand-in that reproduces the *batch contract* of `MDQM9SamplerDataset.process` + PyG collate
(mdqm9/data/mdqm9_ambient.py:160-170, mdqm9/data/mdqm9_latent.py:188-205):

  ambient: x, x0 [N,3] f32 (per-molecule centred), latent_z [N,3], latent_dlogp [B], T0, T1 [N] f32,
           atoms [N] i64 = arange(n) per molecule, edge_index [2,E] i64 (complete digraph,
           (src,dst)-lexicographic after coalesce), edge_type [E] i64 in {0..3}, batch [N] i64, ptr [B+1]
  latent : x = 0, x0 = centred randn, T [N] i64, atom_number [N] i64, same graph fields
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch


class MolBatch:
    """Mutable attribute bag with the few whole-object operations the samplers use
    (`to`, `clone`, `[]`); mirrors torch_geometric.data.Batch for that surface only."""

    def __init__(self, **tensors):
        object.__setattr__(self, "_store", dict(tensors))

    def __getattr__(self, name):
        store = object.__getattribute__(self, "_store")
        try:
            return store[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self._store[name] = value

    def __delattr__(self, name):
        self._store.pop(name, None)

    def __getitem__(self, name):
        return self._store[name]

    def __setitem__(self, name, value):
        self._store[name] = value

    def __contains__(self, name):
        return name in self._store

    def keys(self):
        return list(self._store)

    def to(self, device, non_blocking: bool = False):
        for k, v in list(self._store.items()):
            if torch.is_tensor(v):
                self._store[k] = v.to(device, non_blocking=non_blocking)
        return self

    def pin_memory(self):
        for k, v in list(self._store.items()):
            if torch.is_tensor(v) and not v.is_cuda:
                self._store[k] = v.pin_memory()
        return self

    def clone(self):
        return MolBatch(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._store.items()})

    @property
    def num_graphs(self) -> int:
        return int(self._store["ptr"].numel() - 1)


def complete_digraph(n_atoms: torch.Tensor):
    """edge_index [2,E] (int64) of the complete digraph of every molecule in coalesced
    (src,dst)-lexicographic order, plus node ptr [B+1] and edge ptr [B+1] - vectorised, on
    `n_atoms.device`.  Row of edge (i -> j) in molecule m: edge_ptr[m] + i*(n_m-1) + j - (j>i)."""
    n_atoms = n_atoms.to(torch.long)
    dev = n_atoms.device
    ptr = torch.zeros(n_atoms.numel() + 1, dtype=torch.long, device=dev)
    ptr[1:] = torch.cumsum(n_atoms, 0)
    ne = n_atoms * (n_atoms - 1)
    eptr = torch.zeros_like(ptr)
    eptr[1:] = torch.cumsum(ne, 0)
    n_edges = int(eptr[-1].item())
    mol = torch.repeat_interleave(torch.arange(n_atoms.numel(), device=dev), ne, output_size=n_edges)
    local = torch.arange(n_edges, device=dev) - eptr[mol]
    nm1 = (n_atoms - 1)[mol]
    i = torch.div(local, nm1, rounding_mode="floor")
    k = local - i * nm1
    j = k + (k >= i).to(torch.long)
    edge_index = torch.stack([ptr[mol] + i, ptr[mol] + j], dim=0)
    return edge_index, ptr, eptr


def chain_bond_orders(n: int) -> torch.Tensor:
    """Synthetic bond table used by every synthetic batch: the chain 0-1-...-(n-1), bond orders
    cycling 1,2,1,3.  Returns the dense [n,n] edge-type matrix (0 = radius-graph edge only), i.e. what
    AddRadiusGraph + AddBondGraph + Coalesce(reduce="max") produce (mdqm9/thermo/utils.py:69-125)."""
    m = torch.zeros(n, n, dtype=torch.long)
    for a in range(n - 1):
        m[a, a + 1] = m[a + 1, a] = (1, 2, 1, 3)[a % 4]
    return m


def _edge_types(n_list: Sequence[int]) -> torch.Tensor:
    out = []
    cache = {}
    for n in n_list:
        if n not in cache:
            m = chain_bond_orders(n)
            mask = ~torch.eye(n, dtype=torch.bool)
            cache[n] = m[mask]          # row-major over (i,j), i != j  == (src,dst)-lexicographic
        out.append(cache[n])
    return torch.cat(out)


def _centred_randn(n_list, sigma, gen):
    n_tot = int(sum(n_list))
    x = torch.randn(n_tot, 3, generator=gen) * sigma
    n_atoms = torch.tensor(list(n_list))
    mol = torch.repeat_interleave(torch.arange(len(n_list)), n_atoms)
    mean = torch.zeros(len(n_list), 3).index_add_(0, mol, x) / n_atoms[:, None].to(x.dtype)
    return x - mean[mol], mol


def synthetic_ambient_batch(n_mol: int, n_atoms=9, *, T0: float = 1000.0, T1: float = 300.0,
                            sigma: float = 0.3, seed: int = 0) -> MolBatch:
    """Synthetic batch in the ambient contract (cfg 2).  `n_atoms` is an int or a
    per-molecule sequence."""
    n_list = [n_atoms] * n_mol if isinstance(n_atoms, int) else list(n_atoms)
    assert len(n_list) == n_mol
    gen = torch.Generator().manual_seed(seed)
    x, mol = _centred_randn(n_list, sigma, gen)
    n_t = torch.tensor(n_list)
    edge_index, ptr, _ = complete_digraph(n_t)
    atoms = torch.cat([torch.arange(n) for n in n_list])
    N = x.shape[0]
    return MolBatch(
        x=x.clone(), x0=x.clone(), latent_z=torch.zeros_like(x), latent_dlogp=torch.zeros(n_mol),
        T0=torch.full((N,), float(T0)), T1=torch.full((N,), float(T1)), atoms=atoms,
        edge_index=edge_index, edge_type=_edge_types(n_list), batch=mol, ptr=ptr)


def synthetic_latent_batch(n_mol: int, n_atoms=9, *, T: Optional[int] = 800, seed: int = 0) -> MolBatch:
    """Synthetic batch in the latent contract (noise -> data; mdqm9/data/mdqm9_latent.py:181-205)."""
    n_list = [n_atoms] * n_mol if isinstance(n_atoms, int) else list(n_atoms)
    gen = torch.Generator().manual_seed(seed)
    x0, mol = _centred_randn(n_list, 1.0, gen)
    n_t = torch.tensor(n_list)
    edge_index, ptr, _ = complete_digraph(n_t)
    atom_number = torch.cat([torch.arange(n) for n in n_list])
    N = x0.shape[0]
    fields = dict(x=torch.zeros_like(x0), x0=x0, atom_number=atom_number, edge_index=edge_index,
                  edge_type=_edge_types(n_list), batch=mol, ptr=ptr)
    if T is not None:
        fields["T"] = torch.full((N,), int(T), dtype=torch.long)
    return MolBatch(**fields)


def synthetic_train_batches(n_mol: int, n_atoms=9, seed: int = 0, T0: float = 1000.0, T1: float = 300.0):
    """(batch0, batch1) in the training contract of MDQM9MultiTempDataset.process (mdqm9/data/mdqm9_ambient.py:87-107):
    x [N,3] centred per molecule, T [N] (the temperature repeated per atom), atoms, the coalesced complete digraph."""
    out = []
    for i, T in enumerate((T0, T1)):
        mb = synthetic_ambient_batch(n_mol, n_atoms, seed=seed + i)
        fields = {k: mb[k] for k in ("x", "atoms", "edge_index", "edge_type", "batch", "ptr")}
        fields["T"] = torch.full((mb.x.shape[0],), float(T))
        out.append(MolBatch(**fields))
    return out
