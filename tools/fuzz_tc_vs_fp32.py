"""Stress test: the tensor-core drift against the fp32 CUDA-core drift of the same library on random batch shapes
(molecule counts, mixed molecule sizes 2..25 atoms (n_types = 25), both variants, repeated calls on a shared workspace).  Looks for
protocol races and edge cases that the fixed-shape tests cannot see; prints one line per case and a summary.
Run on a B200:  python tools/fuzz_tc_vs_fp32.py [n_cases] [seed] [path]
path = fused (default: F = 128, k_message_tc / k_update_tc / k_readout_tc), layered (F = 128 on k_chain_tc, two CTAs per SM),
f256 (F = 256 on k_chain_tc), div (drift + exact divergence: tensor-core tangents vs fp32 dual-number kernels, small batches)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._util import perturb_  # noqa: E402
from thermodynamic_interpolation_b200 import _lib  # noqa: E402
from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch, synthetic_latent_batch  # noqa: E402

DEV = "cuda:0"


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    path = sys.argv[3] if len(sys.argv) > 3 else "fused"
    F = 256 if path == "f256" else 128
    tc_mode = _lib.MATH_F16X3_LAYERED if path == "layered" else _lib.MATH_F16X3_TC
    shrink = {"fused": 1, "layered": 1, "f256": 6, "div": 60}[path]      # keep the slower paths' cases small
    gen = torch.Generator().manual_seed(seed)
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN as A
    from thermodynamic_interpolation_b200.latent.models.cpainn import cPaiNN as L
    torch.manual_seed(seed)
    models = {
        "ambient": perturb_(A(n_features=F, score_layers=3, temp_length=100), seed + 1).eval().to(DEV),
        "latent": perturb_(L(n_features=F, score_layers=3, temp_length=75), seed + 2).eval().to(DEV),
    }
    worst, bad, t_start = 0.0, 0, time.time()
    for case in range(n_cases):
        variant = "ambient" if int(torch.randint(0, 2, (1,), generator=gen)) == 0 else "latent"
        kind = int(torch.randint(0, 4, (1,), generator=gen))
        if kind == 0:      # uniform size
            n_mol = int(torch.randint(1, 3000 // shrink + 2, (1,), generator=gen))
            sizes = int(torch.randint(2, 26, (1,), generator=gen))
        elif kind == 1:    # mixed sizes
            n_mol = int(torch.randint(1, 1500 // shrink + 2, (1,), generator=gen))
            lo = int(torch.randint(2, 12, (1,), generator=gen))
            hi = min(25, lo + int(torch.randint(1, 20, (1,), generator=gen)))
            sizes = torch.randint(lo, hi + 1, (n_mol,), generator=gen).tolist()
        elif kind == 2:    # a few molecules (fewer tiles than SMs)
            n_mol = int(torch.randint(1, 12, (1,), generator=gen))
            sizes = torch.randint(2, 26, (n_mol,), generator=gen).tolist()
        else:              # exactly one row short / one node over the tile boundaries
            n_mol = int(torch.randint(100 // shrink + 1, 800 // shrink + 3, (1,), generator=gen))
            sizes = [9] * n_mol
            sizes[int(torch.randint(0, n_mol, (1,), generator=gen))] = int(torch.randint(2, 26, (1,), generator=gen))
        mk = synthetic_ambient_batch if variant == "ambient" else synthetic_latent_batch
        mb = mk(n_mol, sizes, seed=1000 + case).to(DEV)
        model = models[variant]
        t = float(torch.rand(1, generator=gen))
        model.set_math(tc_mode)
        eng = model.engine()
        pb = eng.prepare(mb)
        if path == "div":
            outs = [eng.drift_div(pb, mb.x0.contiguous(), t)[1].clone() for _ in range(2)]
            eng.status()
            model.set_math(_lib.MATH_FP32_SIMT)
            ref = model.engine().drift_div(pb, mb.x0.contiguous(), t)[1].clone()
            bound = 5e-4      # one scalar per molecule: a cancelling sum of 3 n diagonal Jacobian entries (worst seen: 2.4e-4, one 19-atom molecule)
        else:
            outs = [eng.drift(pb, mb.x0.contiguous(), t).clone() for _ in range(2)]
            eng.status()
            model.set_math(_lib.MATH_FP32_SIMT)
            ref = model.engine().drift(pb, mb.x0.contiguous(), t).clone()
            bound = 5e-5 if F == 128 else 1e-4
        scale = float(ref.abs().max())
        err = float((outs[0] - ref).abs().max()) / scale
        rep = float((outs[0] - outs[1]).abs().max())
        ok = bool(torch.isfinite(outs[0]).all()) and err < bound and rep == 0.0
        worst = max(worst, err)
        bad += 0 if ok else 1
        szs = sizes if isinstance(sizes, int) else f"{min(sizes)}..{max(sizes)}"
        print(f"case {case:3d} {variant:7s} kind {kind} n_mol {n_mol:5d} atoms {szs!s:8s} N {mb.x0.shape[0]:6d} "
              f"err {err:.2e} repeat-diff {rep:.1e} {'ok' if ok else 'FAIL'}", flush=True)
    print(f"path {path}: {n_cases} cases, {bad} failures, worst tensor-core-vs-fp32 error {worst:.2e} of the largest reference value, {time.time() - t_start:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
