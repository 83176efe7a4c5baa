"""Random-shape stress test of the training step (tib_train_loss_grad) against torch autograd through the CPU oracle
(oracle/train_oracle.py, pinned against the unmodified reference):

    python tools/fuzz_train_vs_oracle.py [n_cases] [seed]

Every case draws a feature width (32 / 64 / 128 / 256), a depth (1..3), a gamma, 1..6 molecules of 2..14 atoms (mixed sizes)
and compares the loss and every gradient tensor (error relative to the tensor's own max, floored at 1e-3 of the largest
gradient entry of the model: a 2-element bias gradient that is a cancelling sum of 30 terms is not held to 1e-4 of itself)."""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import train_oracle as to  # noqa: E402
from tests._util import oracle_hp_sd  # noqa: E402
from thermodynamic_interpolation_b200.batch import synthetic_train_batches  # noqa: E402
from thermodynamic_interpolation_b200.engine import packed_keys  # noqa: E402
from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model  # noqa: E402
from thermodynamic_interpolation_b200.train import TrainEngine, flatten, packed_parameters  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    worst_g, worst_l, fails = 0.0, 0.0, 0
    t_start = time.time()
    for case in range(n_cases):
        F = int(rng.choice([32, 64, 128, 128, 256]))
        L = int(rng.integers(1, 4))
        gamma = str(rng.choice(["sin2", "brownian"]))
        sizes = [int(v) for v in rng.integers(2, 15, size=int(rng.integers(1, 7)))]
        model = seeded_ambient_model(F, L, seed=int(rng.integers(0, 1000))).to("cuda:0")
        hp, sd = oracle_hp_sd(model)
        b0, b1 = synthetic_train_batches(len(sizes), sizes, int(rng.integers(0, 1000)))
        torch.manual_seed(int(rng.integers(0, 1 << 30)))
        t, z = to.draw_t_z(sizes)
        t = t.clamp(0.03, 0.97)                     # brownian gamma_dot is singular at the ends
        eng = TrainEngine(model.hyper, "cuda:0")
        loss, grad, _ = eng.loss_and_grad(flatten(packed_parameters(model)), eng.prepare(b0, b1), t, z, gamma=gamma)
        eng.status()
        ref_loss, ref_grads, _, _ = to.loss_and_grads(sd, hp, b0.x, b1.x, t, z, b0.atoms, b0.edge_index, b0.edge_type, b0.T, b1.T,
                                                      gamma=gamma)
        el = abs(float(loss) - float(ref_loss)) / max(1.0, abs(float(ref_loss)))
        eg, off, flat = 0.0, 0, grad.cpu()
        gmax = max(float(v.abs().max()) for v in ref_grads.values() if v is not None)
        for k, shp in packed_keys(model.hyper):
            n = int(np.prod(shp))
            ref = ref_grads[k].reshape(-1)
            e_k = float((flat[off:off + n] - ref).abs().max()) / max(float(ref.abs().max()), 1e-3 * gmax, 1e-30)
            if e_k > 1e-4 and os.environ.get("FUZZ_VERBOSE"):
                print(f"    {k}: err {e_k:.2e}, max |ref| {float(ref.abs().max()):.3e}, |grad| total {float(flat.abs().max()):.3e}")
            eg = max(eg, e_k)
            off += n
        ok = el < 1e-4 and eg < 3e-4 and bool(torch.isfinite(flat).all())
        fails += 0 if ok else 1
        worst_g, worst_l = max(worst_g, eg), max(worst_l, el)
        print(f"case {case:3d} F={F:3d} L={L} gamma={gamma:8s} atoms={sizes}: loss err {el:.1e}, worst gradient tensor err {eg:.1e} {'ok' if ok else 'FAIL'}",
              flush=True)
    print(f"{n_cases} cases, {fails} failures, worst loss error {worst_l:.2e}, worst gradient-tensor error {worst_g:.2e} "
          f"({time.time() - t_start:.0f} s)")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
