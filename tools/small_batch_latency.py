"""Latency of the sampler at the reference's real batch sizes (config/ambient/00031_settings_no_300.json:18 batch_size 12,
10506_settings_no_900.json:18 batch_size 64) and at 256 / 4096: per Euler step, eager launches vs one CUDA graph per rollout.
Run on a B200:  python tools/small_batch_latency.py > profiles/r02_small_batch_latency.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thermodynamic_interpolation_b200 import _lib  # noqa: E402
from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch  # noqa: E402
from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model  # noqa: E402

DEV = "cuda:0"


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    K = 50
    grid = torch.linspace(0.0, 1.0, K + 1)
    for F, n_atoms, sizes in ((128, 9, (12, 64, 256, 4096)), (256, 25, (12, 64, 256))):
        model = seeded_ambient_model(F, 5, 100).to(DEV)
        eng = model.engine()
        for B in sizes:
            mb = synthetic_ambient_batch(B, n_atoms, seed=100).to(DEV)
            pb = eng.prepare(mb)
            x0 = mb.x0.contiguous()
            out = torch.empty((K + 1, x0.shape[0], 3), device=DEV)
            eager = lambda: eng.rollout_fixed(pb, x0, grid, method="euler", save_frames=True, out=out)  # noqa: E731
            graph = lambda: eng.rollout_fixed(pb, x0, grid, method="euler", save_frames=True, out=out, graph=True)  # noqa: E731
            eager(); ref = out.clone(); graph()
            same = bool(torch.equal(ref, out))
            ms_e = timed(eager, 5) / K
            ms_g = timed(graph, 5) / K
            eng.status()
            print(json.dumps(dict(workload=f"{B} conformers x {n_atoms} atoms, F={F} L=5, {K} Euler steps", math=_lib.MATH_NAMES[1],
                                  ms_per_step_eager=ms_e, ms_per_step_cuda_graph=ms_g, speedup=ms_e / ms_g,
                                  mol_steps_per_s_cuda_graph=B / (ms_g * 1e-3), graph_bit_identical=same)), flush=True)


if __name__ == "__main__":
    main()
