"""CPU emulation of the split-f16 GEMM arithmetic (csrc/tc_common.cuh) inside the oracle, against an fp64 evaluation of the
same network: which operand scales and which of the three passes (a_hi*b_hi, a_hi*b_lo, a_lo*b_hi; a = activations, b =
weights) each GEMM class needs to stay at fp32-level error - the per-GEMM pass ablation of VERDICT r1 item 3(iii).

    python tools/emulate_split_f16.py [--steps 100] [--mols 4]  > profiles/r02_split_f16_pass_ablation.txt

Emulated: the rounding of both operands to hi + lo f16 (with the operand scale), the dropped passes; products and sums in
fp64 (the tensor core's fp32 accumulation is not modelled, so these are lower bounds of the device error)."""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpainn_oracle as co  # noqa: E402
from tests._util import oracle_hp_sd, perturb_  # noqa: E402
from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN  # noqa: E402
from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch  # noqa: E402

F_ = torch.nn.functional
CLASSES = ["phi1", "phi2", "phi3", "w1", "w2", "w3", "upd1", "upd2", "upd3", "uv", "ro1", "ro2"]


class Emu:
    def __init__(self, state_scale=2.0 ** -4, hidden_scale=1.0, passes=None):
        self.state_scale, self.hidden_scale = state_scale, hidden_scale
        self.passes = {c: "3" for c in CLASSES}
        self.passes.update(passes or {})

    def linear(self, x, W, b, scale, cls):
        p = self.passes[cls]
        if scale == "row":
            m = x.abs().amax(dim=-1, keepdim=True).clamp_min(1e-30)
            sc = torch.exp2(9.0 - torch.floor(torch.log2(m)))
        else:
            sc = torch.tensor(float(scale))
        xs = (x * sc).float()
        xh = xs.half().float()
        xl = (xs - xh).half().float()
        wh = W.half().float()
        wl = (W - wh).half().float()
        acc = xh.double() @ wh.double().T
        if p in ("3", "2w"):
            acc = acc + xh.double() @ wl.double().T
        if p in ("3", "2a"):
            acc = acc + xl.double() @ wh.double().T
        y = (acc / sc.double()).float()
        return y + b if b is not None else y


def patched(emu, L):
    """(mlp, equivariant_linear) replacements for the oracle module."""
    def cls_of(prefix):
        if ".phi." in prefix:
            return "phi"
        if ".w." in prefix:
            return "w"
        idx = int(prefix.split("layers.")[1].split(".")[0]) if "layers." in prefix else -1
        if idx == 2 * L:
            return "ro"
        return "upd" if idx >= 0 else None

    def mlp(x, sd, prefix):
        c = cls_of(prefix)
        if c is None:      # embedding MLP: fp32 CUDA cores in the product
            return ORIG_MLP(x, sd, prefix)
        in_scale = 1.0 if c == "w" else emu.state_scale      # w's input is the positional encoding
        h = emu.linear(x, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"], in_scale, c + "1")
        h = F_.silu(F_.layer_norm(h, (h.shape[-1],), sd[f"{prefix}.1.weight"], sd[f"{prefix}.1.bias"], 1e-5))
        h = emu.linear(h, sd[f"{prefix}.3.weight"], sd[f"{prefix}.3.bias"], emu.hidden_scale, c + "2")
        h = F_.silu(F_.layer_norm(h, (h.shape[-1],), sd[f"{prefix}.4.weight"], sd[f"{prefix}.4.bias"], 1e-5))
        if c == "ro":      # the readout's last Linear is a dot product in fp32
            return F_.linear(h, sd[f"{prefix}.6.weight"], sd[f"{prefix}.6.bias"])
        return emu.linear(h, sd[f"{prefix}.6.weight"], sd[f"{prefix}.6.bias"], emu.hidden_scale, c + "3")

    def eq_linear(weight, v):
        if weight.shape[0] == 1:     # readout V: fp32 dot product
            return ORIG_EQ(weight, v)
        return emu.linear(v.swapaxes(-1, -2), weight, None, emu.state_scale, "uv").swapaxes(-1, -2)

    return mlp, eq_linear


ORIG_MLP, ORIG_EQ = co.mlp, co.equivariant_linear


def run(emu, sd, hp, mb, x, t):
    if emu is not None:
        co.mlp, co.equivariant_linear = patched(emu, hp.score_layers)
    try:
        with torch.no_grad():
            return co.drift(sd, hp, x, t, mb.atoms, mb.edge_index, mb.edge_type, T0=mb.T0.to(x.dtype), T1=mb.T1.to(x.dtype))
    finally:
        co.mlp, co.equivariant_linear = ORIG_MLP, ORIG_EQ


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mols", type=int, default=4)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--drift-mols", type=int, default=32)
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    model = perturb_(cPaiNN(n_features=128, score_layers=5, temp_length=100), 1).eval()
    hp, sd = oracle_hp_sd(model)
    sd64 = {k: v.double() for k, v in sd.items()}
    mb = synthetic_ambient_batch(args.drift_mols, 9, T0=1000.0, T1=300.0, sigma=0.3, seed=100)
    truth = run(None, sd64, hp, mb, mb.x0.double(), 0.3)
    fp32 = run(None, sd, hp, mb, mb.x0, 0.3)
    print(f"network: cfg 2 (F=128, L=5, 9 atoms), {args.drift_mols} conformers, t=0.3; error = max|b - b64| / max|b64|")
    print(f"{'fp32 oracle (the reference arithmetic)':64s} {rel(fp32, truth):.3e}")
    configs = [
        ("x3, state 2^-4, hidden 1 (round-1 kernels)", Emu()),
        ("x3, state 2^-4, hidden 2^8", Emu(hidden_scale=256.0)),
        ("x3, per-row scales everywhere", Emu(state_scale="row", hidden_scale="row")),
        ("x3, state per-row, hidden 2^8", Emu(state_scale="row", hidden_scale=256.0)),
    ]
    for c in CLASSES:
        configs.append((f"2 passes on {c} only: weights hi-only (a_hi*b_hi + a_lo*b_hi)", Emu(hidden_scale=256.0, passes={c: "2a"})))
    for c in CLASSES:
        configs.append((f"2 passes on {c} only: activations hi-only (a_hi*b_hi + a_hi*b_lo)", Emu(hidden_scale=256.0, passes={c: "2w"})))
    configs.append(("2 passes everywhere: weights hi-only", Emu(hidden_scale=256.0, passes={c: "2a" for c in CLASSES})))
    configs.append(("2 passes everywhere: activations hi-only", Emu(hidden_scale=256.0, passes={c: "2w" for c in CLASSES})))
    configs.append(("1 pass everywhere (plain f16)", Emu(hidden_scale=256.0, passes={c: "1" for c in CLASSES})))
    for name, emu in configs:
        print(f"{name:64s} {rel(run(emu, sd, hp, mb, mb.x0, 0.3), truth):.3e}", flush=True)

    if args.steps > 0:
        # trajectories: fixed-grid Euler, every arithmetic from the same x0; error vs the fp64 trajectory
        mbt = synthetic_ambient_batch(args.mols, 9, T0=1000.0, T1=300.0, sigma=0.3, seed=100)
        K = args.steps
        dt = 1.0 / K
        marks = sorted({1, 10, 50, K} & set(range(1, K + 1)))
        trajs = {"fp64": (None, sd64, mbt.x0.double()), "fp32 oracle": (None, sd, mbt.x0.clone()),
                 "x3 round-1 scales": (Emu(), sd, mbt.x0.clone()), "x3 hidden 2^8": (Emu(hidden_scale=256.0), sd, mbt.x0.clone()),
                 "x3 per-row scales": (Emu(state_scale="row", hidden_scale="row"), sd, mbt.x0.clone()),
                 "2 passes everywhere, weights hi-only": (Emu(hidden_scale=256.0, passes={c: "2a" for c in CLASSES}), sd, mbt.x0.clone())}
        out = {k: {} for k in trajs}
        state = {k: v[2] for k, v in trajs.items()}
        for k in range(1, K + 1):
            for name, (emu, sdd, _) in trajs.items():
                x = state[name]
                b = run(emu, sdd, hp, mbt, x, (k - 1) * dt)
                state[name] = x + dt * b
            if k in marks:
                for name in trajs:
                    out[name][k] = rel(state[name], state["fp64"])
        print(f"\nEuler, {K} steps, {args.mols} conformers: max|x_k - x_k(fp64)| / max|x_k(fp64)| at k = {marks}")
        for name in trajs:
            if name != "fp64":
                print(f"{name:64s} " + "  ".join(f"{out[name][k]:.3e}" for k in marks))


if __name__ == "__main__":
    main()
