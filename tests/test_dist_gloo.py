"""CPU tier: the N>1 host path under torch.distributed (gloo, world_size 2): shard, gather, all-reduce."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cpainn_oracle as co
from thermodynamic_interpolation_b200 import batch as B, dist as D, stats as S


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = B.synthetic_ambient_batch(7, [9, 12, 9, 25, 10, 9, 9], seed=5)
        mine = D.shard_batch(full, rank, world)
        # stand-in for the per-rank rollout result: a deterministic function of the shard's x0
        local_final = mine.x0 * 2.0 + 1.0
        gathered = D.gather_samples(local_final)
        ok_gather = torch.equal(gathered, full.x0 * 2.0 + 1.0)
        # reweighting partial sums: per-rank (numpy stand-in for the CUDA kernel), one fp64 all-reduce
        rng = np.random.default_rng(3)
        E0, E1, nd = rng.normal(0, 1, 1000), rng.normal(0.2, 1, 1000), rng.normal(0, 0.3, 1000)
        lo, hi = D.shard_range(1000, rank, world)
        w = co.ti_weights(E0[lo:hi], E1[lo:hi], nd[lo:hi])
        part = torch.tensor([w.sum(), (w * w).sum(), w.sum(), float(hi - lo), float(hi - lo)], dtype=torch.float64)
        tot = S.finalize(D.allreduce_stats(part))
        w_all = co.ti_weights(E0, E1, nd)
        ok_stats = (abs(tot["ess"] - co.ess(w_all)) < 1e-9 * co.ess(w_all) and tot["n"] == 1000 and
                    abs(tot["dF"] - co.tfep_dF(E1 - E0 + nd, np.ones(1000))) < 1e-12)
        # dopri5 error-norm callback: {sum_sq, count} summed over ranks
        cb = D.norm_allreduce(torch.device("cpu"))
        import ctypes as C
        buf = (C.c_double * 2)(float(rank + 1), 10.0)
        cb(buf, None)
        ok_norm = (buf[0], buf[1]) == (3.0, 20.0)
        ret[rank] = (ok_gather, ok_stats, ok_norm)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: (True, True, True), 1: (True, True, True)}
