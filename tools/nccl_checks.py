"""Collectives of the sampler on hardware (NCCL over NVLink), world size >= 2:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_checks.py

 1. dopri5 with the batch-global error norm shared through `norm_allreduce` (one 2-double all-reduce per attempted step):
    every shard takes the SAME step sequence as the unsharded run - equal NFE / attempts, frames equal to the unsharded
    frames of its molecules (torchdiffeq's RMS norm couples the whole batch).
 2. reweighting partial sums + one fp64 all-reduce == the unsharded statistics; all-gather of the final samples.
 3. the IQR outlier mask from GLOBAL percentiles (sensititvity.py:4-12) == the mask computed on the gathered vector.
 4. data-parallel training (train_ambient.Trainer(data_parallel=True)): every rank computes the gradient of its own batch,
    ONE NCCL all-reduce averages the flat gradient vector, every rank applies the same clipped Adam step - the weights after
    two steps equal (a) each other on all ranks and (b) a single-process run that averages the per-rank gradients itself.
Exit code 0 = all checks passed on every rank."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thermodynamic_interpolation_b200 import _lib, analysis as A, dist as D, stats as S  # noqa: E402
from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator  # noqa: E402
from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch  # noqa: E402
from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    model = seeded_ambient_model(128, 3, 100).to(dev)
    model.set_math(_lib.MATH_FP32_SIMT)            # op-by-op fp32: shard and full batch evaluate identical per-molecule arithmetic
    full = synthetic_ambient_batch(64, 9, seed=5)
    mine = D.shard_batch(full, rank, world)
    lo, hi = D.shard_range(64, rank, world)
    # ---- 1. dopri5
    integ_full = MoleculeIntegrator(model, method="dopri5", n_step=6, atol=1e-5, rtol=1e-5)
    x_full, _, nfe_full, _ = integ_full.rollout(full.clone().to(dev))
    integ = MoleculeIntegrator(model, method="dopri5", n_step=6, atol=1e-5, rtol=1e-5, norm_allreduce=D.norm_allreduce(dev))
    x_mine, _, nfe, _ = integ.rollout(mine.clone().to(dev))
    n_lo, n_hi = int(full.ptr[lo]), int(full.ptr[hi])
    err = float((x_mine - x_full[:, n_lo:n_hi]).abs().max() / x_full.abs().max())
    same = nfe == nfe_full and integ.last_stats["attempts"] == integ_full.last_stats["attempts"]
    print(f"[nccl rank {rank}] dopri5 shared norm: nfe {nfe} (unsharded {nfe_full}), frames vs unsharded {err:.2e}", flush=True)
    ok &= same and err < 1e-5
    # without the shared norm the shards are free to take different step sequences (documented behaviour)
    integ_local = MoleculeIntegrator(model, method="dopri5", n_step=6, atol=1e-5, rtol=1e-5)
    _, _, nfe_local, _ = integ_local.rollout(mine.clone().to(dev))
    print(f"[nccl rank {rank}] dopri5 local norm: nfe {nfe_local}", flush=True)
    # ---- 2. statistics + gather
    gen = torch.Generator().manual_seed(9)
    E0 = torch.randn(64, generator=gen, dtype=torch.float64)
    E1 = E0 + 0.3 * torch.randn(64, generator=gen, dtype=torch.float64)
    part = S.reweight_partials(E0[lo:hi].to(dev), E1[lo:hi].to(dev))
    tot = S.finalize(D.allreduce_stats(part).cpu())
    ref = S.finalize(S.reweight_partials(E0.to(dev), E1.to(dev)).cpu())
    ok &= abs(tot["ess"] - ref["ess"]) < 1e-9 * ref["ess"] and abs(tot["dF"] - ref["dF"]) < 1e-12 and tot["n"] == 64
    final = x_mine[-1].reshape(hi - lo, 9, 3)
    gathered = D.gather_samples(final)
    ok &= tuple(gathered.shape) == (64, 9, 3) and torch.equal(gathered[lo:hi], final)
    # ---- 3. IQR mask with global percentiles
    w = torch.exp(-(E1 - E0))
    w[3] = 1e6                                      # an outlier on rank 0's shard
    keep_local = A.filter_iqr(w[lo:hi].to(dev), k=3)
    q25, q75 = torch.quantile(w, torch.tensor([0.25, 0.75], dtype=torch.float64)).tolist()
    keep_ref = (w > q25 - 3 * (q75 - q25)) & (w < q75 + 3 * (q75 - q25))
    ok &= torch.equal(keep_local.cpu(), keep_ref[lo:hi]) and not bool(keep_ref[3])
    # ---- 4. data-parallel training step
    from thermodynamic_interpolation_b200.ambient.interpolants import LinearInterpolant
    from thermodynamic_interpolation_b200.batch import synthetic_train_batches
    from thermodynamic_interpolation_b200.train_ambient import Trainer
    ip = LinearInterpolant(a=1, gamma="sin2")

    def draws(r, step, n_mol=16):
        g = torch.Generator().manual_seed(100 * r + step)
        return torch.rand(n_mol, generator=g).repeat_interleave(9).reshape(-1, 1), torch.randn(n_mol * 9, 3, generator=g)

    batches = [synthetic_train_batches(16, 9, seed=50 + r) for r in range(world)]
    dp = Trainer(seeded_ambient_model(128, 2, 100, seed=3).to(dev), ip, data_parallel=True)
    solo = Trainer(seeded_ambient_model(128, 2, 100, seed=3).to(dev), ip)
    err_g = 0.0
    for step in range(2):
        t, z = draws(rank, step)
        dp.step(*batches[rank], t=t, z=z)
        grads = []
        for r in range(world):                       # the same thing without NCCL: every rank's gradient, averaged by hand
            t, z = draws(r, step)
            grads.append(solo.loss_and_grad(*batches[r], t=t, z=z)[1])
        mean_g = torch.stack(grads).mean(0)
        if step == 0:                                # identical weights so far: the all-reduced gradient must equal the hand average
            err_g = float((dp.last_grad - mean_g).abs().max() / mean_g.abs().max())
        solo.apply(mean_g)
    gathered_w = [torch.empty_like(dp.weights) for _ in range(world)]
    dist.all_gather(gathered_w, dp.weights)
    same_w = all(torch.equal(gathered_w[0], w_) for w_ in gathered_w)
    err_w = float((dp.weights - solo.weights).abs().max())
    print(f"[nccl rank {rank}] data-parallel training: replicas identical {same_w}, all-reduced gradient vs hand average {err_g:.2e}, "
          f"weights after 2 steps {err_w:.2e}", flush=True)
    # Adam normalises every element by its own magnitude, so elements whose gradient is at rounding level (the split-K sums are
    # atomic: their order differs from run to run) may move by up to lr per step in either run: the weight bound is 2 lr + slack
    ok &= same_w and err_g < 1e-5 and err_w < 2.5e-4
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("NCCL CHECKS", "PASSED" if int(flag) else "FAILED", f"(world size {world})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
