"""Drop-in for the driver loop of mdqm9/sample_latent.py (`sample`, lines 18-100): batches -> latent rollout ->
the reference's on-disk format

    samples_<name>_forward.npy   [n_mol, T, n_atoms, 3]   frames per molecule
    dlogps_<name>_forward.npy    [n_mol]                  last-frame dlogp (only with return_dlogp)

which mdqm9/data/mdqm9_ambient.py:173-199 chains into the ambient flow (it reads frames [:, 0] and [:, -1]).
As in sample_ambient.py the dataset is out of scope (`loader` is any iterable of batches in the latent batch
contract), frames come back through one pinned D2H copy, and the arrays are written every `save_every`
batches instead of after every batch.  The reference also drops a second copy of both arrays into the current
working directory at the end (sample_latent.py:88-92); that is kept behind `also_cwd=True`."""
from __future__ import annotations

import argparse
import os
from typing import Iterable, Optional

import numpy as np
import torch

from .latent.integrators import MoleculeIntegrator
from .sample_ambient import regroup_frames


def sample(config: argparse.Namespace, b: torch.nn.Module, loader: Iterable, *, method: str = "dopri5",
           device: Optional[str] = None, save_every: int = 1, also_cwd: bool = False, verbose: bool = True) -> dict:
    """Reads from `config`: seed, data_save_path, data_save_name, rtol, atol, n_steps, return_dlogp."""
    torch.manual_seed(config.seed)
    np.random.seed(config.seed)
    os.makedirs(config.data_save_path, exist_ok=True)
    dev = torch.device(device or "cuda")
    integrator = MoleculeIntegrator(b=b, method=method, rtol=config.rtol, atol=config.atol, n_step=config.n_steps,
                                    return_dlogp=bool(config.return_dlogp), reverse_ode=False)
    b.eval()
    b.to(dev)
    samples, dlogps = [], []
    stem = lambda s, root: os.path.join(root, f"{s}_{config.data_save_name}_forward.npy")  # noqa: E731

    def flush(root):
        if samples:
            np.save(stem("samples", root), np.concatenate(samples, axis=0))
        if config.return_dlogp and dlogps:
            np.save(stem("dlogps", root), np.concatenate(dlogps, axis=0))

    host_frames = None
    n_batches = 0
    for i, batch in enumerate(loader):
        batch = batch.to(dev)
        xts, dlogp, batch_idx = integrator.rollout(batch)
        if host_frames is None or host_frames.shape != xts.shape:
            host_frames = torch.empty(xts.shape, dtype=xts.dtype).pin_memory()
        host_frames.copy_(xts, non_blocking=True)
        torch.cuda.synchronize(dev)
        samples.append(regroup_frames(host_frames.numpy(), batch_idx.detach().cpu().numpy()))
        if config.return_dlogp:
            dlogps.append(dlogp.detach().cpu().numpy()[-1, :])                  # sample_latent.py:76
        n_batches = i + 1
        if save_every and n_batches % save_every == 0:
            flush(config.data_save_path)
        if verbose:
            print(f"Batch {n_batches}")
    flush(config.data_save_path)
    if also_cwd:
        flush(os.getcwd())
    if verbose:
        print("Finished forward sampling...\n")
    return dict(samples=np.concatenate(samples, axis=0) if samples else None,
                dlogps=np.concatenate(dlogps, axis=0) if dlogps else None)
