"""Drop-in for the driver loop of mdqm9/sample_ambient.py (`sample`, lines 18-119): batches -> rollout ->
the reference's on-disk format, so its analysis scripts (mdqm9/analysis/results_00031.py:173-201) and the
latent->ambient chaining (mdqm9/data/mdqm9_ambient.py:173-199) read our output unchanged:

    samples_<name>.npy        [n_mol, T, n_atoms, 3]   frames per molecule
    dlogps_<name>.npy         [n_mol]                  last-frame dlogp (only with return_dlogp)
    latent_noises_<name>.npy  [n_mol, n_atoms, 3],  latent_dlogps_<name>.npy [n_mol]

The dataset (RDKit + trajectory files, mdqm9/data/mdqm9_ambient.py) is out of scope: `loader` is any iterable of
batches in the batch contract (thermodynamic_interpolation_b200.batch).  Differences to the reference loop, none
visible in the files: frames come back through one pinned-memory D2H copy and a vectorised regroup instead of a
Python loop over molecules, and the arrays are written once per `save_every` batches (reference: the whole
accumulated array after every batch - O(batches^2) I/O)."""
from __future__ import annotations

import argparse
import os
from typing import Iterable, Optional

import numpy as np
import torch

from .ambient.integrators import MoleculeIntegrator


def regroup_frames(xts: np.ndarray, batch_idx: np.ndarray) -> np.ndarray:
    """[T, N, 3] node-major frames -> [B, T, n, 3] per-molecule trajectories; the reference's
    `np.array([sample[:, batch_idx == i] for i in range(batch_idx.max() + 1)])` (sample_ambient.py:93).
    Equal molecule sizes (the reference's own assumption: np.array of ragged pieces would not be a
    rectangular array) and sorted `batch_idx` are required."""
    n_mol = int(batch_idx.max()) + 1
    counts = np.bincount(batch_idx, minlength=n_mol)
    if counts.min() != counts.max():
        raise ValueError("regroup_frames needs equal molecule sizes (as the reference's np.array(...) does)")
    if np.any(np.diff(batch_idx) < 0):
        raise ValueError("batch_idx must be sorted")
    n = int(counts[0])
    T = xts.shape[0]
    return np.ascontiguousarray(xts.reshape(T, n_mol, n, xts.shape[-1]).transpose(1, 0, 2, 3))


def sample(config: argparse.Namespace, b: torch.nn.Module, loader: Iterable, *, method: str = "dopri5",
           device: Optional[str] = None, save_every: int = 1, verbose: bool = True) -> dict:
    """Runs every batch of `loader` through `MoleculeIntegrator.rollout` and writes the reference's files
    under config.data_save_path.  Reads from `config`: seed, data_save_path, data_save_name, rtol, atol,
    n_steps, return_dlogp.  Returns the arrays it wrote."""
    torch.manual_seed(config.seed)
    np.random.seed(config.seed)
    os.makedirs(config.data_save_path, exist_ok=True)
    dev = torch.device(device or "cuda")
    integrator = MoleculeIntegrator(b=b, method=method, rtol=config.rtol, atol=config.atol, n_step=config.n_steps,
                                    return_dlogp=bool(config.return_dlogp), reverse_ode=False)
    b.eval()
    b.to(dev)
    out = dict(latent_noises=[], latent_dlogps=[], samples=[], dlogps=[])
    path = lambda stem: os.path.join(config.data_save_path, f"{stem}_{config.data_save_name}.npy")  # noqa: E731

    def flush():
        for stem in ("latent_noises", "latent_dlogps", "samples") + (("dlogps",) if config.return_dlogp else ()):
            if out[stem]:
                np.save(path(stem), np.concatenate(out[stem], axis=0))

    n_steps = None
    host_frames = None
    for i, batch in enumerate(loader):
        batch = batch.to(dev)
        batch_idx = batch.batch.detach().cpu().numpy()
        if hasattr(batch, "latent_z") or "latent_z" in getattr(batch, "keys", lambda: [])():
            out["latent_noises"].append(regroup_frames(batch.latent_z.detach().cpu().numpy()[None], batch_idx)[:, 0])
            out["latent_dlogps"].append(batch.latent_dlogp.detach().cpu().numpy())
        xts, dlogp, n_steps, _ = integrator.rollout(batch)
        if host_frames is None or host_frames.shape != xts.shape:
            host_frames = torch.empty(xts.shape, dtype=xts.dtype).pin_memory()
        host_frames.copy_(xts, non_blocking=True)
        torch.cuda.synchronize(dev)
        out["samples"].append(regroup_frames(host_frames.numpy(), batch_idx))
        if config.return_dlogp:
            out["dlogps"].append(dlogp.detach().cpu().numpy()[-1, :])
        if save_every and (i + 1) % save_every == 0:
            flush()
        if verbose:
            print(f"Batch {i + 1}")
    if verbose:
        print(f"Number sampling steps: {n_steps}")
    flush()
    return {k: (np.concatenate(v, axis=0) if v else None) for k, v in out.items()}
