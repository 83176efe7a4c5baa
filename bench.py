#!/usr/bin/env python
"""bench.py - molecule*SDE-steps/sec of the MDQM9 ambient sampler hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg4|cfg5]

One bench "step" = one integrator step of the sampler over the whole batch: one drift-network evaluation
b(t, x, T0, T1) on every molecule plus the fused state update (the unit SURVEY.md section 8d defines).

Workloads (BASELINE.json configs):
  cfg2 (default, configs[1]): 4096 conformers x 9 atoms per GPU, cPaiNN F=128 L=5, T0=1000 K -> T1=300 K, fixed-grid Euler
        on linspace(0, 1, K+1), all K+1 frames saved.  Weak scaling: every rank integrates its own 4096 conformers; no
        collective on the data path.
  cfg4 (configs[3]): the temperature sweep - 125 000 conformers per GPU (1 M on 8 GPUs) in chunks of 15 625, T1 cycling over
        {300, ..., 900} K (mdqm9/config/ambient/00031_settings_no_{300..900}.json:32-33), K Euler steps per temperature, final
        samples only; after every temperature the reweighting statistics (tib_reweight_stats -> ONE fp64 all-reduce of 5
        numbers) and the IQR outlier mask from the all-gathered weights (mdqm9/analysis/utils/sensititvity.py:4-12, k = 100 as
        at results_00031.py:248-268) - the collectives are INSIDE the timed region.

  cfg5 (configs[4]): the training step of train_ambient.py - per GPU 256 molecules x 9 atoms (two batches at T0 / T1), loss =
        StandardVelocityLoss(LinearInterpolant(a=1, gamma='sin2')) with both antithetic drift evaluations, backward, one NCCL
        all-reduce of the gradient vector (data parallel, averaged over ranks), clip_grad_norm_(1), Adam(lr=1e-4).  One bench
        step = one optimisation step; the unit is molecule*training-steps/s (molecules of batch0 per second).

`value`    : device-resident run (inputs already in HBM), CUDA events, max over ranks.
`e2e`      : the same work through the public API `MoleculeIntegrator.rollout(batch)` with the batch in pinned host memory
             (H2D inside the timed region) and the result read back to pinned host memory (D2H).
`roofline` : dominant kernel (the fused message kernel, 30 F^2 FLOP per edge per layer) from CUDA event pairs recorded around
             each of its launches in the timed region.
`cpu_baseline` / `--impl reference`: the reference's OWN modules (staged under baseline/_ref by tools/stage_reference.py,
             imported behind oracle/stubs) running MoleculeIntegrator(method='euler').rollout on the host cores, B_cpu = 256
             (BASELINE.md section 5); when the staged files are absent, the oracle port (kind "port").
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
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
SWEEP_T1 = [300.0, 400.0, 500.0, 600.0, 700.0, 800.0, 900.0]


def flops_per_mol_step(n, F, L):
    """SURVEY.md section 8d (2 x MACs; biases, LN, SiLU, trig ignored)."""
    e = n * (n - 1)
    return n * 12 * F * F + L * (e * 30 * F * F + n * 24 * F * F) + n * (4 * F * F + 4 * F + 6 * F)


def message_flops_per_launch(n_mol, n, F):
    return n_mol * n * (n - 1) * 30 * F * F


def kernel_source_sha():
    h = hashlib.sha256()
    for f in ("tc_message.cuh", "tc_common.cuh"):
        with open(os.path.join(REPO, "thermodynamic_interpolation_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the ncu capture of THIS build of the kernel
    (profiles/r02_ncu_traffic.json carries the hash of the kernel sources it was taken from); None otherwise."""
    p = os.path.join(REPO, "profiles", "r02_ncu_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)["k_message_tc"]
        if d.get("src_sha") != kernel_source_sha():
            return None
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


def build_model_and_batch(args, seed, n_mol, T1=300.0):
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model
    model = seeded_ambient_model(args.features, args.layers, 100, seed=0)
    batch = synthetic_ambient_batch(n_mol, args.atoms, T0=1000.0, T1=T1, sigma=0.3, seed=100 + seed)
    return model, batch


def workload_name(args):
    if args.workload == "cfg4":
        return (f"MDQM9 ambient temperature sweep, {args.mols} conformers x {args.atoms} atoms per GPU in chunks of {args.chunk}, "
                f"T1 in {{300..900}} K, cPaiNN F={args.features} L={args.layers}, fixed-grid Euler, per-temperature reweighting "
                f"all-reduce + IQR mask (BASELINE configs[3])")
    return (f"MDQM9 ambient sampling, {args.mols} conformers x {args.atoms} atoms per GPU, cPaiNN F={args.features} "
            f"L={args.layers}, fixed-grid Euler (BASELINE configs[1])")


# ---------------------------------------------------------------------------------------------------
# CPU arms (oracle/ and baseline/_ref are used here and only here)
# ---------------------------------------------------------------------------------------------------
def cpu_port_rate(args, n_mol, budget_s, min_steps, fixed_steps=None, warmup=1):
    """The oracle PORT (oracle/cpainn_oracle.py) Euler steps on the host cores: (mol*steps/s, cores, steps, seconds)."""
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


def reference_root():
    """The staged copy of the reference's own modules (GPU box) or the reference itself (build container)."""
    for p in (os.path.join(REPO, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "mdqm9", "thermo")):
            return p
    return None


def cpu_reference_rate(args, n_mol, steps, warmup):
    """The UNMODIFIED reference: MoleculeIntegrator(b, method='euler', n_step=steps + 1).rollout(batch) on the host cores
    (mdqm9/thermo/ambient/integrators.py:28-68), weights from the same seeded recipe as the B200 arm.
    Returns (mol*steps/s, cores, seconds) or None when the reference files are not available."""
    root = reference_root()
    if root is None:
        return None
    os.environ["TI_REFERENCE_ROOT"] = root
    from oracle import ref_loader
    from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch
    from thermodynamic_interpolation_b200.synthetic import perturb_
    ref_loader.REFERENCE_ROOT = root
    ns = ref_loader.load_mdqm9()
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = perturb_(ns.ambient_cpainn.cPaiNN(n_features=args.features, score_layers=args.layers, temp_length=100), 1).eval()
    # TemperatureEncoder parks its `temperatures` attribute on cuda whenever a GPU is visible (embedding.py:204); this is the
    # CPU arm, so the attribute follows the parameters (what `.to(device)` does for registered tensors)
    for mod in model.modules():
        if isinstance(getattr(mod, "temperatures", None), torch.Tensor):
            mod.temperatures = mod.temperatures.cpu()
    mb = synthetic_ambient_batch(n_mol, args.atoms, T0=1000.0, T1=300.0, sigma=0.3, seed=100)

    def ref_batch():
        b = ns.torch_geometric.data.Batch()
        for k in mb.keys():
            b[k] = mb[k].clone() if torch.is_tensor(mb[k]) else mb[k]
        return b

    if warmup > 0:
        ns.ambient_integrators.MoleculeIntegrator(model, method="euler", n_step=warmup + 1).rollout(ref_batch())
    integ = ns.ambient_integrators.MoleculeIntegrator(model, method="euler", n_step=steps + 1)
    b = ref_batch()
    t0 = time.perf_counter()
    xts, dlogp, nfe, bvec = integ.rollout(b)
    el = time.perf_counter() - t0
    assert tuple(xts.shape) == (steps + 1, mb.x0.shape[0], 3) and torch.isfinite(xts).all()
    return n_mol * steps / el, cores, el


def run_reference(args):
    """`--impl reference`: rank 0 only; the other ranks exit without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu = min(args.mols, 256)
    ref = cpu_reference_rate(args, n_cpu, args.steps, args.warmup)
    if ref is not None:
        rate, cores, secs = ref
        kind = "reference"
        sample = (f"the reference's MoleculeIntegrator(method='euler', n_step={args.steps + 1}).rollout on {n_cpu} conformers "
                  f"of the same workload ({secs:.1f} s, {cores} threads); the sweep / statistics part of cfg4 is not included")
        port_rate, _, port_steps, port_secs = cpu_port_rate(args, 128, budget_s=4.0, min_steps=2)
        port = dict(value=port_rate, unit=UNIT, kind="port", sample=f"oracle port, 128 conformers x {port_steps} Euler steps ({port_secs:.1f} s)")
    else:
        n_cpu = min(args.mols, 128)
        rate, cores, steps, secs = cpu_port_rate(args, n_cpu, budget_s=0.0, min_steps=args.steps, fixed_steps=args.steps, warmup=args.warmup)
        kind = "port"
        sample = f"oracle port (reference files not staged): {n_cpu} conformers x {steps} Euler steps of the same workload ({secs:.1f} s)"
        port = None
    line = dict(metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * secs / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=workload_name(args), sample=sample),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=cores, kind=kind, sample=sample),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    if port:
        line["cpu_baseline_port"] = port
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# cfg5: the training step (BASELINE configs[4])
# ---------------------------------------------------------------------------------------------------
TRAIN_METRIC = "molecule*training-steps/sec, MDQM9 ambient drift network (loss fwd+bwd, clip, Adam)"


def train_workload_name(args):
    return (f"MDQM9 ambient training step, {args.mols} molecules x {args.atoms} atoms per GPU (batch0 at 1000 K, batch1 at 300 K), "
            f"cPaiNN F={args.features} L={args.layers}, StandardVelocityLoss(LinearInterpolant(a=1, gamma='sin2')), antithetic "
            f"pair, clip_grad_norm_(1), Adam(lr=1e-4), data-parallel gradient all-reduce (BASELINE configs[4])")


def train_flops_per_mol(n, F, L):
    """Algorithmic FLOPs of one training step per molecule: two antithetic drift evaluations, each forward + data gradient +
    weight gradient (3 x the forward contractions of SURVEY.md section 8d)."""
    return 2 * 3 * flops_per_mol_step(n, F, L)


def cpu_reference_train_rate(args, n_mol, steps, warmup):
    """The UNMODIFIED reference training step on the host cores (mdqm9/train_ambient.py:124-148): loss_fn(batch0, batch1, model),
    backward, clip_grad_norm_(1), Adam.  Returns (mol*steps/s, cores, seconds) or None."""
    root = reference_root()
    if root is None:
        return None
    os.environ["TI_REFERENCE_ROOT"] = root
    from oracle import ref_loader
    from thermodynamic_interpolation_b200.batch import synthetic_train_batches
    from thermodynamic_interpolation_b200.synthetic import perturb_
    ref_loader.REFERENCE_ROOT = root
    ns = ref_loader.load_mdqm9()
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = perturb_(ns.ambient_cpainn.cPaiNN(n_features=args.features, score_layers=args.layers, temp_length=100), 1).train()
    for mod in model.modules():
        if isinstance(getattr(mod, "temperatures", None), torch.Tensor):
            mod.temperatures = mod.temperatures.cpu()
    b0, b1 = synthetic_train_batches(n_mol, args.atoms, seed=100)

    def ref_batch(mb):
        b = ns.torch_geometric.data.Batch()
        for k in mb.keys():
            b[k] = mb[k].clone() if torch.is_tensor(mb[k]) else mb[k]
        return b

    loss_fn = ns.ambient_losses.StandardVelocityLoss(interpolant=ns.ambient_interpolants.LinearInterpolant(a=1, gamma="sin2"),
                                                     t_distr="uniform")
    optim = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=0)

    def one():
        optim.zero_grad()
        loss = loss_fn(ref_batch(b0), ref_batch(b1), model)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        optim.step()
        return float(loss)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        last = one()
    el = time.perf_counter() - t0
    assert np.isfinite(last)
    return n_mol * steps / el, cores, el


def run_reference_train(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu = min(args.mols, 64)
    steps = max(1, min(args.steps, 5))
    ref = cpu_reference_train_rate(args, n_cpu, steps, min(args.warmup, 1))
    if ref is None:
        print(json.dumps(dict(impl="reference", unavailable="the reference's training modules are not staged under baseline/_ref")))
        return
    rate, cores, secs = ref
    sample = (f"the reference's own training step (StandardVelocityLoss -> backward -> clip_grad_norm_ -> Adam) on {n_cpu} molecules, "
              f"{steps} steps ({secs:.1f} s, {cores} threads)")
    print(json.dumps(dict(metric=TRAIN_METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                          ms_per_step=1e3 * secs / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                          data="synthetic", impl="reference", config=dict(workload=train_workload_name(args), sample=sample),
                          cpu_baseline=dict(value=rate, unit=UNIT, cores=cores, kind="reference", sample=sample),
                          e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)))


def run_train(args):
    import torch.distributed as dist
    from thermodynamic_interpolation_b200 import _lib
    from thermodynamic_interpolation_b200.ambient.interpolants import LinearInterpolant
    from thermodynamic_interpolation_b200.batch import synthetic_train_batches
    from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model
    from thermodynamic_interpolation_b200.train_ambient import Trainer

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
    model = seeded_ambient_model(args.features, args.layers, 100, seed=0).to(dev)
    trainer = Trainer(model, LinearInterpolant(a=1, gamma="sin2"), lr=1e-4, weight_decay=0.0, max_grad_norm=1.0,
                      data_parallel=world > 1, cuda_graph=bool(args.graph))
    n_host = 4                                        # distinct host batches cycled through (each rank its own data)
    host = []
    gen = torch.Generator().manual_seed(1234 + rank)
    for i in range(n_host):
        b0, b1 = synthetic_train_batches(args.mols, args.atoms, seed=1000 * rank + 10 * i)
        N = b0.x.shape[0]
        t = torch.rand(args.mols, generator=gen).repeat_interleave(args.atoms).reshape(N, 1)
        z = torch.randn(N, 3, generator=gen)
        host.append((b0.pin_memory(), b1.pin_memory(), t.pin_memory(), z.pin_memory()))
    devb = [(b0.clone().to(dev), b1.clone().to(dev), t.to(dev), z.to(dev)) for b0, b1, t, z in host]
    tbs = [trainer.engine.prepare(b0, b1) for b0, b1, _, _ in devb]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        tt = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    losses = torch.zeros(2 * K + W + 16, dtype=torch.float64, device=dev)

    def resident_step(i):
        tb, (_, _, t, z) = tbs[i % n_host], devb[i % n_host]
        loss, grad = trainer.loss_and_grad(None, None, t=t, z=z, prepared=tb)
        trainer.apply(grad)
        losses[i] = loss[0]

    ms_sum = (C.c_double * _lib.N_KERNEL_KINDS)()
    launches = (C.c_uint64 * _lib.N_KERNEL_KINDS)()
    step_no = 0
    use_graph = bool(args.graph)
    trainer.cuda_graph = False
    for _ in range(W):                                         # W untimed warm-up steps (eager launches)
        resident_step(step_no); step_no += 1
    barrier()
    profiled_steps, ms_profiled, n_launch_p, gemm_flops_p = 0, 0.0, 0, 0.0
    if use_graph:
        # a graph replay cannot carry per-launch event pairs (and issues no launches of its own): the kernel-class times of
        # the roofline come from eager launches of the same kernels on the same inputs, right before the timed region
        profiled_steps = min(K, 5)
        lib.tib_launch_count(1)
        lib.tib_train_gemm_flops(1)
        lib.tib_profile_begin()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(profiled_steps):
            resident_step(step_no); step_no += 1
        p1.record()
        barrier()
        n_launch_p = int(lib.tib_launch_count(0))
        gemm_flops_p = float(lib.tib_train_gemm_flops(0))
        _lib.check(lib.tib_profile_end(ms_sum, launches), "tib_profile_end")
        ms_profiled = p0.elapsed_time(p1)
        trainer.cuda_graph = True
        resident_step(step_no); step_no += 1                   # captures the graph (once per batch shape) + one replay
        barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    lib.tib_launch_count(1)
    lib.tib_train_gemm_flops(1)
    if not use_graph:
        lib.tib_profile_begin()
    t_begin = time.perf_counter()
    e0.record()
    for _ in range(K):
        resident_step(step_no); step_no += 1
    e1.record()
    barrier()
    t_end = time.perf_counter()
    if use_graph:
        n_launch = n_launch_p * K // profiled_steps            # kernels inside the K replayed steps
        gemm_flops = gemm_flops_p * K / profiled_steps
    else:
        n_launch = int(lib.tib_launch_count(0))
        gemm_flops = float(lib.tib_train_gemm_flops(0))
        _lib.check(lib.tib_profile_end(ms_sum, launches), "tib_profile_end")
        profiled_steps = K
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    if not use_graph:
        ms_profiled = ms_total
    clk = clocks.stop(t_begin, t_end) if clocks else None
    trainer.engine.status()
    lh = losses[:step_no].cpu()
    assert torch.isfinite(lh).all(), "training diverged"
    value = world * args.mols * K / (ms_total * 1e-3)

    # ---- e2e: host batches (pinned) -> H2D, prepare, step, loss read back - every step
    host_loss = torch.zeros(1, dtype=torch.float64).pin_memory()
    h2d = sum(v.numel() * v.element_size() for b in host[0][:2] for v in (b[k] for k in b.keys()) if torch.is_tensor(v))
    h2d += host[0][2].numel() * 4 + host[0][3].numel() * 4

    def e2e_step(i):
        b0, b1, t, z = host[i % n_host]
        loss = trainer.step(b0, b1, t=t.to(dev, non_blocking=True), z=z.to(dev, non_blocking=True))     # pinned host batches
        host_loss.copy_(loss, non_blocking=True)

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * args.mols * K / e2e_s

    if rank == 0:
        peaks = load_peaks()
        gi = _lib.KERNEL_KINDS.index("train_gemm")
        gemm_ms, gemm_n = ms_sum[gi], int(launches[gi])
        other_ms = ms_sum[_lib.KERNEL_KINDS.index("train_other")]
        achieved = (gemm_flops * profiled_steps / K) / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        peak = peaks["bf16_sustained"]
        fl = train_flops_per_mol(args.atoms, args.features, args.layers)
        line = dict(
            metric=TRAIN_METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total / K,
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=train_workload_name(args), mols_per_gpu=args.mols, atoms=args.atoms, n_features=args.features,
                        layers=args.layers, math="f16x3_tcgen05", optimizer="Adam(lr=1e-4), clip_grad_norm_(1)",
                        cuda_graph=bool(args.graph),
                        l2="activations saved for the backward pass (%.0f MB per step) exceed the 126 MB L2" %
                           (lib.tib_train_workspace_bytes(C.byref(trainer.engine.desc), tbs[0].pb.n_mol, tbs[0].pb.n_nodes, tbs[0].pb.n_edges) / 1e6),
                        flops_per_mol_step=fl, whole_step_tflops=value * fl / 1e12 / world,
                        loss_first=float(lh[0]), loss_last=float(lh[-1])),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=8, seconds=e2e_s,
                     api="train_ambient.Trainer.step(host batch0, host batch1) + D2H of the loss"),
            gpu_launches=n_launch, clocks=clk,
            roofline=dict(kernel="k_gemm_tc (every Linear of the step: forward, data gradient, weight gradient)", bound="tensor",
                          achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=None,
                          peak_source=peaks["source"] + ", bf16 dense sustained", launches=gemm_n,
                          avg_launch_ms=gemm_ms / max(gemm_n, 1), flops_per_step=gemm_flops / K,
                          kernel_time_shares=dict(train_gemm=round(gemm_ms / ms_profiled, 4), train_other=round(other_ms / ms_profiled, 4)),
                          measured_on=(f"{profiled_steps} eager steps right before the timed region ({ms_profiled / profiled_steps:.2f} ms per step; graph "
                                       "replays carry no per-launch events)" if args.graph else "the timed region"),
                          note="two streams: the kernel-class shares add up to more than the wall time they overlap in"))
        if world == 1 and not args.no_cpu:
            ref = cpu_reference_train_rate(args, min(args.mols, 64), steps=2, warmup=1)
            if ref is not None:
                rate, cores, secs = ref
                line["cpu_baseline"] = dict(value=rate, unit=UNIT, cores=cores, kind="reference",
                                            sample=f"the reference's own training step, {min(args.mols, 64)} molecules x 2 steps ({secs:.1f} s)")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    sweep = args.workload == "cfg4"
    temps = SWEEP_T1 if sweep else [300.0]
    chunk = min(args.chunk, args.mols) if sweep else args.mols
    n_chunks = (args.mols + chunk - 1) // chunk

    model, _ = build_model_and_batch(args, rank, 1)
    model = model.to(dev)
    model.set_math(args.math)
    eng = model.engine()
    # host chunks (pinned), device copies and prepared batches per (chunk, temperature is stamped into T1 in place)
    host_chunks, dev_chunks = [], []
    for c in range(n_chunks):
        n_c = min(chunk, args.mols - c * chunk)
        _, hb = build_model_and_batch(args, rank * 1000 + c, n_c)
        hb.pin_memory()
        host_chunks.append(hb)
        dev_chunks.append(hb.clone().to(dev))
    grid = torch.linspace(0.0, 1.0, K + 1)
    n_nodes0 = dev_chunks[0].x0.shape[0]
    frames = None if sweep else torch.empty((K + 1, n_nodes0, 3), dtype=torch.float32, device=dev)
    finals = [torch.empty_like(db.x0) for db in dev_chunks]

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

    def energies(x0, x1, n_mol, T1):
        """Synthetic reduced energies (SURVEY.md section 8d cfg 4): harmonic E = 1/2 |x|^2 (1000 / T)."""
        E0 = 0.5 * (x0.reshape(n_mol, -1) ** 2).sum(1).double()
        E1 = 0.5 * (x1.reshape(n_mol, -1) ** 2).sum(1).double() * (1000.0 / T1)
        return E0, E1

    stats_out = {}

    def sweep_statistics(T1):
        """Per temperature: partial sums on every rank -> one fp64 all-reduce; IQR mask on the all-gathered weights."""
        part = torch.zeros(S.N_STATS, dtype=torch.float64, device=dev)
        ws = []
        for db, xf in zip(dev_chunks, finals):
            n_c = int(db.ptr.numel() - 1)
            E0, E1 = energies(db.x0, xf, n_c, T1)
            shift = 0.0                                            # weights are used up to a constant; keep exp() in range
            part += S.reweight_partials(E0, E1 - shift)
            ws.append(torch.exp(-(E1 - E0)))
        tot = S.finalize(D.allreduce_stats(part).cpu())
        w_all = D.gather_samples(torch.cat(ws))                    # [world * mols] fp64: the filter needs GLOBAL percentiles
        q25, q75 = torch.quantile(w_all, torch.tensor([0.25, 0.75], dtype=w_all.dtype, device=dev)).tolist()
        iqr = q75 - q25
        keep = (w_all > q25 - 100 * iqr) & (w_all < q75 + 100 * iqr)          # filter_iqr(k=100)
        stats_out[T1] = dict(ess=tot["ess"], dF=tot["dF"], n=tot["n"], kept=int(keep.sum().item()))

    def device_pass(pbs_by_T):
        for T1 in temps:
            for c, db in enumerate(dev_chunks):
                eng.rollout_fixed(pbs_by_T[T1][c], db.x0, grid, method="euler", save_frames=not sweep,
                                  out=frames if not sweep else finals[c])
            if sweep:
                sweep_statistics(T1)

    # prepared batches per temperature (the sweep re-stamps T1; the x-independent embedding tables depend on it)
    pbs_by_T = {}
    for T1 in temps:
        pbs = []
        for db in dev_chunks:
            db.T1 = torch.full_like(db.T1, T1)          # a new tensor per temperature: prepared batches keep what they were made from
            pbs.append(eng.prepare(db))
        pbs_by_T[T1] = pbs

    # ---- warm-up: W untimed steps of the same hot path
    if W > 0:
        eng.rollout_fixed(pbs_by_T[temps[0]][0], dev_chunks[0].x0, grid[: W + 1], method="euler", save_frames=False)
    barrier()

    # ---- timed region: exactly K steps per (temperature, conformer), inputs resident in HBM
    clocks = ClockSampler(local) if rank == 0 else None
    ms_sum = (C.c_double * _lib.N_KERNEL_KINDS)()
    launches = (C.c_uint64 * _lib.N_KERNEL_KINDS)()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    lib.tib_launch_count(1)
    lib.tib_profile_begin()
    t_begin = time.perf_counter()
    e0.record()
    device_pass(pbs_by_T)
    e1.record()
    barrier()
    t_end = time.perf_counter()
    n_launch = int(lib.tib_launch_count(0))
    _lib.check(lib.tib_profile_end(ms_sum, launches), "tib_profile_end")
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop(t_begin, t_end) if clocks else None
    eng.status()
    last = frames[-1] if not sweep else finals[-1]
    assert torch.isfinite(last).all(), "rollout diverged"
    units = args.mols * K * len(temps)
    value = world * units / (ms_total * 1e-3)

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region
    integ = MoleculeIntegrator(model, method="euler", n_step=K + 1, save_frames=not sweep)
    out_shape = (K + 1, n_nodes0, 3) if not sweep else (n_nodes0, 3)
    host_out = torch.empty(out_shape, dtype=torch.float32).pin_memory()
    h2d = sum(sum(v.numel() * v.element_size() for v in (hb[k] for k in hb.keys()) if torch.is_tensor(v)) for hb in host_chunks) * len(temps)
    d2h = host_out.numel() * 4 * n_chunks * len(temps)

    def e2e_once():
        for T1 in temps:
            for c, hb in enumerate(host_chunks):
                b = hb.clone().to(dev, non_blocking=True)
                b.T1.fill_(T1)
                xts, dlogp, nfe, bvec = integ.rollout(b)
                if xts.shape == host_out.shape:
                    host_out.copy_(xts, non_blocking=True)
                else:
                    host_out[: xts.shape[0]].copy_(xts, non_blocking=True)
                if sweep:
                    finals[c][: xts.shape[0]].copy_(xts)
            if sweep:
                sweep_statistics(T1)
        torch.cuda.synchronize(dev)

    if not sweep:
        e2e_once()                   # warm (allocator, workspace)
    barrier()
    t0 = time.perf_counter()
    e2e_once()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * units / e2e_s

    if not sweep:
        # the only collectives of the cfg-2 job: final sample gather + statistics all-reduce (outside the step loop)
        db = dev_chunks[0]
        E0, E1 = energies(db.x0, frames[-1], args.mols, 300.0)
        part = S.reweight_partials(E0 - E0.mean(), E1 - E1.mean())
        tot = S.finalize(D.allreduce_stats(part).cpu())
        stats_out[300.0] = dict(ess=tot["ess"], dF=tot["dF"], n=tot["n"])
        if world > 1 and args.gather:
            gathered = D.gather_samples(frames[-1].reshape(args.mols, args.atoms, 3))
            assert gathered.shape[0] == world * args.mols
    barrier()

    if rank == 0:
        peaks = load_peaks()
        msg_ms = ms_sum[_lib.KERNEL_KINDS.index("message")]
        msg_n = int(launches[_lib.KERNEL_KINDS.index("message")])
        per_launch_ms = msg_ms / max(msg_n, 1)
        mflops = message_flops_per_launch(chunk, args.atoms, args.features)
        achieved = mflops / (per_launch_ms * 1e-3) / 1e12
        peak = peaks["bf16_sustained"]
        shares = {k: round(ms_sum[i] / ms_total, 4) for i, k in enumerate(_lib.KERNEL_KINDS)}
        e_bytes = pbs_by_T[temps[0]][0].n_edges * args.features * 4
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total / (K * len(temps) * n_chunks),
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=workload_name(args), mols_per_gpu=args.mols, atoms=args.atoms, n_features=args.features,
                        layers=args.layers, method="euler", math=_lib.MATH_NAMES[args.math],
                        frames_saved=(K + 1) if not sweep else 1, temperatures=temps, chunk=chunk,
                        l2="per-step working set (edge features e[E,F] = %.0f MB + node features) exceeds the 126 MB L2" % (e_bytes / 1e6),
                        flops_per_mol_step=flops_per_mol_step(args.atoms, args.features, args.layers),
                        whole_step_tflops=value * flops_per_mol_step(args.atoms, args.features, args.layers) / 1e12 / world),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d / (K * len(temps) * n_chunks),
                     d2h_bytes_per_step=d2h / (K * len(temps) * n_chunks), seconds=e2e_s,
                     api="ambient.integrators.MoleculeIntegrator.rollout(host batch) + D2H of " + ("all frames" if not sweep else "the final samples")),
            gpu_launches=n_launch, clocks=clk,
            roofline=dict(kernel=("k_message_tc" if args.math != 0 else "k_message") + " (SE3Message: phi/w edge MLPs + gated scatter)", bound="tensor",
                          achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                          traffic=load_traffic() if (args.math == 1 and chunk == 4096 and args.atoms == 9) else None,
                          peak_source=peaks["source"] + ", bf16 dense sustained", launches=msg_n,
                          avg_launch_ms=per_launch_ms, flops_per_launch=mflops, kernel_time_shares=shares),
            stats={str(int(T)): v for T, v in stats_out.items()})
        if world == 1 and not args.no_cpu:
            ref = cpu_reference_rate(args, 256, steps=max(2, min(K, 8)), warmup=1)
            if ref is not None:
                rate, cores, secs = ref
                line["cpu_baseline"] = dict(value=rate, unit=UNIT, cores=cores, kind="reference",
                                            sample=f"the reference's MoleculeIntegrator(method='euler').rollout, 256 conformers x {max(2, min(K, 8))} steps ({secs:.1f} s)")
            else:
                rate, cores, steps, secs = cpu_port_rate(args, 128, budget_s=args.cpu_seconds, min_steps=2)
                line["cpu_baseline"] = dict(value=rate, unit=UNIT, cores=cores, kind="port",
                                            sample=f"128 conformers x {steps} Euler steps of the same workload ({secs:.1f} s)")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4", "cfg5"])
    ap.add_argument("--mols", type=int, default=None, help="conformers per GPU (cfg2: 4096, cfg4: 125000)")
    ap.add_argument("--chunk", type=int, default=15625, help="cfg4: conformers per rollout call")
    ap.add_argument("--atoms", type=int, default=9)
    ap.add_argument("--features", type=int, default=128)
    ap.add_argument("--layers", type=int, default=5)
    ap.add_argument("--math", type=int, default=1, help="TIB_MATH_*: 0 fp32 SIMT, 1 split-f16 x3 tcgen05 (default), 2 single-pass f16")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--gather", type=int, default=1)
    ap.add_argument("--graph", type=int, default=1, help="cfg5: replay the captured launches of the loss / gradient call (CUDA graph)")
    args = ap.parse_args()
    if args.mols is None:
        args.mols = {"cfg4": 125000, "cfg5": 256}.get(args.workload, 4096)
    if args.steps is None:
        args.steps = {"cfg4": 20, "cfg5": 20}.get(args.workload, 200)
    if args.workload == "cfg5":
        (run_reference_train if args.impl == "reference" else run_train)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
