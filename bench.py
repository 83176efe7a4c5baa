#!/usr/bin/env python
"""bench.py - molecule*SDE-steps/sec of the MDQM9 ambient sampler hot path (BASELINE.json cfg 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One bench "step" = one integrator step of the sampler over the whole batch: one drift-network
evaluation b(t, x, T0, T1) on every molecule plus the fused state update and frame write (the unit
SURVEY.md section 8d defines).  Workload (configs[1]): 4096 conformers x 9 atoms per GPU, cPaiNN
F=128, L=5, T0=1000 K -> T1=300 K, fixed-grid Euler on linspace(0, 1, K+1), random-init weights,
synthetic centred coordinates.  Weak scaling: every rank integrates its own 4096 conformers; there
is no collective on the data path (trajectories are independent) - the final statistics all-reduce
is outside the per-step loop and is exercised once after the timed region.

`value`    : device-resident rollout (inputs already in HBM), CUDA events, max over ranks.
`e2e`      : the same K steps through the public API `MoleculeIntegrator.rollout(batch)` with the
             batch in pinned host memory (H2D inside the timed region) and all K+1 frames read back
             to pinned host memory (D2H), as mdqm9/sample_ambient.py:74,88-91 does.
`roofline` : dominant kernel (the per-molecule message kernel, 30 F^2 FLOP per edge per layer) from
             CUDA event pairs recorded around each of its launches in the timed region.
`cpu_baseline`: the CPU oracle (restatement of the reference, validated against the unmodified
             reference's outputs in tests/golden) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "molecule*SDE-steps/sec, MDQM9 ambient sampler"
UNIT = "molecule*steps/s"


def flops_per_mol_step(n, F, L):
    """SURVEY.md section 8d (2 x MACs; biases, LN, SiLU, trig ignored)."""
    e = n * (n - 1)
    return n * 12 * F * F + L * (e * 30 * F * F + n * 24 * F * F) + n * (4 * F * F + 4 * F + 6 * F)


def message_flops_per_launch(n_mol, n, F):
    return n_mol * n * (n - 1) * 30 * F * F


def load_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture (None if absent)."""
    p = os.path.join(REPO, "profiles", "r01_ncu_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)["k_message_tc"]
        return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:
        return None


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows = []
        self.proc = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            sel = ["-i", uuid if uuid.startswith("GPU-") else "GPU-" + uuid]
        except Exception:
            sel = ["-i", str(device_index)]
        try:
            self.proc = subprocess.Popen(["nvidia-smi", *sel, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [r for (ts, r) in self.rows if t_begin <= ts <= t_end + 0.2] or [r for (_, r) in self.rows]
        for r in rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                    samples=len(sm), reasons=sorted(reasons))


def build_model_and_batch(args, seed, n_mol):
    from tests._util import perturb_
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    torch.manual_seed(0)
    model = perturb_(cPaiNN(n_features=args.features, score_layers=args.layers, temp_length=100), 1).eval()
    batch = synthetic_ambient_batch(n_mol, args.atoms, T0=1000.0, T1=300.0, sigma=0.3, seed=100 + seed)
    return model, batch


def workload_name(args):
    return (f"MDQM9 ambient sampling, {args.mols} conformers x {args.atoms} atoms per GPU, cPaiNN F={args.features} "
            f"L={args.layers}, fixed-grid Euler (BASELINE configs[1])")


# ---------------------------------------------------------------------------------------------------
def cpu_oracle_rate(args, n_mol, budget_s, min_steps, fixed_steps=None, warmup=1):
    """Oracle Euler steps on the host cores: returns (mol*steps/s, cores, steps, seconds)."""
    from oracle import cpainn_oracle as co
    from tests._util import oracle_hp_sd
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    model, mb = build_model_and_batch(args, 0, n_mol)
    hp, sd = oracle_hp_sd(model)
    x = mb.x0.clone()
    dt = 1.0 / 200

    def step(x, k):
        with torch.no_grad():
            b = co.drift(sd, hp, x, k * dt, mb.atoms, mb.edge_index, mb.edge_type, T0=mb.T0, T1=mb.T1)
        return x + dt * b

    for k in range(warmup):
        x = step(x, k)
    t0 = time.perf_counter()
    done = 0
    while True:
        x = step(x, done)
        done += 1
        el = time.perf_counter() - t0
        if fixed_steps is not None:
            if done >= fixed_steps:
                break
        elif done >= min_steps and el >= budget_s:
            break
    el = time.perf_counter() - t0
    assert torch.isfinite(x).all()
    return n_mol * done / el, cores, done, el


def run_reference(args):
    """`--impl reference`: the reference algorithm's CPU implementation (oracle port - the reference
    itself needs torch_geometric / torch_scatter / torchdiffeq, absent from this image and the GPU
    box) on all host threads, same workload shape, each step a bounded sample of 128 conformers."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu = min(args.mols, 128)
    rate, cores, steps, secs = cpu_oracle_rate(args, n_cpu, budget_s=0.0, min_steps=args.steps,
                                               fixed_steps=args.steps, warmup=args.warmup)
    sample = f"{n_cpu} conformers x {steps} Euler steps of the same workload ({secs:.1f} s)"
    line = dict(metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * secs / steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=workload_name(args), sample=sample),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from thermodynamic_interpolation_b200 import _lib
    from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    from thermodynamic_interpolation_b200 import dist as D, stats as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    K, W = args.steps, max(args.warmup, 0)

    model, host_batch = build_model_and_batch(args, rank, args.mols)
    host_batch.pin_memory()
    model = model.to(dev)
    model.set_math(args.math)
    eng = model.engine()
    dbatch = host_batch.clone().to(dev)
    pb = eng.prepare(dbatch)
    x0 = dbatch.x0.contiguous()
    grid = torch.linspace(0.0, 1.0, K + 1)
    frames = torch.empty((K + 1, pb.n_nodes, 3), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up: W untimed steps of the same hot path
    if W > 0:
        eng.rollout_fixed(pb, x0, torch.linspace(0.0, 1.0, K + 1)[: W + 1], method="euler", save_frames=False)
    barrier()

    # ---- timed region: exactly K steps, inputs resident in HBM
    clocks = ClockSampler(local) if rank == 0 else None
    ms_sum = (C.c_double * _lib.N_KERNEL_KINDS)()
    launches = (C.c_uint64 * _lib.N_KERNEL_KINDS)()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    lib.tib_launch_count(1)
    lib.tib_profile_begin()
    t_begin = time.perf_counter()
    e0.record()
    eng.rollout_fixed(pb, x0, grid, method="euler", save_frames=True, out=frames)
    e1.record()
    barrier()
    t_end = time.perf_counter()
    n_launch = int(lib.tib_launch_count(0))
    _lib.check(lib.tib_profile_end(ms_sum, launches), "tib_profile_end")
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop(t_begin, t_end) if clocks else None
    assert torch.isfinite(frames[-1]).all(), "rollout diverged"
    value = world * args.mols * K / (ms_total * 1e-3)

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region
    integ = MoleculeIntegrator(model, method="euler", n_step=K + 1)
    host_out = torch.empty((K + 1, pb.n_nodes, 3), dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in (host_batch[k] for k in host_batch.keys()) if torch.is_tensor(v))
    d2h = host_out.numel() * 4

    def e2e_once():
        b = host_batch.clone().to(dev, non_blocking=True)
        xts, dlogp, nfe, bvec = integ.rollout(b)
        host_out.copy_(xts, non_blocking=True)
        torch.cuda.synchronize(dev)

    e2e_once()                       # warm (allocator, workspace)
    barrier()
    t0 = time.perf_counter()
    e2e_once()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * args.mols * K / e2e_s

    # ---- the only collectives of the job: final sample gather + statistics all-reduce (outside the step loop)
    final = frames[-1].reshape(args.mols, args.atoms, 3)
    E0 = 0.5 * (x0.reshape(args.mols, -1) ** 2).sum(1).double() * (1000.0 / 1000.0)
    E1 = 0.5 * (final.reshape(args.mols, -1) ** 2).sum(1).double() * (1000.0 / 300.0)
    part = S.reweight_partials(E0 - E0.mean(), E1 - E1.mean())
    tot = S.finalize(D.allreduce_stats(part).cpu())
    if world > 1 and args.gather:
        gathered = D.gather_samples(final)
        assert gathered.shape[0] == world * args.mols
    barrier()

    if rank == 0:
        peaks = load_peaks()
        msg_ms = ms_sum[_lib.KERNEL_KINDS.index("message")]
        msg_n = int(launches[_lib.KERNEL_KINDS.index("message")])
        per_launch_ms = msg_ms / max(msg_n, 1)
        mflops = message_flops_per_launch(args.mols, args.atoms, args.features)
        achieved = mflops / (per_launch_ms * 1e-3) / 1e12
        peak = peaks["bf16_sustained"]
        shares = {k: round(ms_sum[i] / ms_total, 4) for i, k in enumerate(_lib.KERNEL_KINDS)}
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total / K,
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=workload_name(args), mols_per_gpu=args.mols, atoms=args.atoms, n_features=args.features,
                        layers=args.layers, method="euler", math=_lib.MATH_NAMES[args.math], frames_saved=K + 1,
                        l2="per-step working set (edge features e[E,F] = %.0f MB + node features) exceeds the 126 MB L2"
                           % (pb.n_edges * args.features * 4 / 1e6),
                        flops_per_mol_step=flops_per_mol_step(args.atoms, args.features, args.layers),
                        whole_step_tflops=value * flops_per_mol_step(args.atoms, args.features, args.layers) / 1e12 / world),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d / K, d2h_bytes_per_step=d2h / K,
                     seconds=e2e_s, api="ambient.integrators.MoleculeIntegrator.rollout(host batch) + D2H of all frames"),
            gpu_launches=n_launch, clocks=clk,
            roofline=dict(kernel=("k_message_tc" if args.math != 0 else "k_message") + " (SE3Message: phi/w edge MLPs + gated scatter)", bound="tensor",
                          achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                          traffic=load_traffic() if args.math != 0 else None,
                          peak_source=peaks["source"] + ", bf16 dense sustained", launches=msg_n,
                          avg_launch_ms=per_launch_ms, flops_per_launch=mflops, kernel_time_shares=shares),
            stats=dict(ess=tot["ess"], dF=tot["dF"], n=tot["n"]))
        if world == 1 and not args.no_cpu:
            rate, cores, steps, secs = cpu_oracle_rate(args, 128, budget_s=args.cpu_seconds, min_steps=2)
            line["cpu_baseline"] = dict(value=rate, unit=UNIT, cores=cores, kind="port",
                                        sample=f"128 conformers x {steps} Euler steps of the same workload ({secs:.1f} s)")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mols", type=int, default=4096, help="conformers per GPU")
    ap.add_argument("--atoms", type=int, default=9)
    ap.add_argument("--features", type=int, default=128)
    ap.add_argument("--layers", type=int, default=5)
    ap.add_argument("--math", type=int, default=1, help="TIB_MATH_*: 0 fp32 SIMT, 1 split-f16 x3 tcgen05 (default), 2 single-pass f16")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--gather", type=int, default=1)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
