"""Additional measurements next to bench.py's headline (BASELINE.json configs 2-4), one JSON line each:
  * cfg 3: latent multi-T flow, 16 384 molecules with 9..25 atoms, F = 128, fixed-grid Euler
  * cfg 2 with the reference's own solver: dopri5, rtol = atol = 1e-5, 100 saved frames (NFE-based throughput)
  * 10506-shaped batch (25 atoms, F = 256): the fp32 SIMT path (tensor cores are built for F = 128)
  * cfg 2 with return_dlogp=True: the drift plus its exact divergence (27 forward-mode tangent directions),
    with the CPU oracle's autograd divergence timed on a 32-conformer sample beside it
  * cfg 1 (ADW): FCNetMultiBeta fp64, 10 000 samples, 100 Euler steps with the exact 1-D divergence, and the CPU
    oracle (the reference's own loop restated) on the same workload
Run on a B200:  python tools/extra_benchmarks.py > gpurun_out/extra.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._util import perturb_  # noqa: E402
from thermodynamic_interpolation_b200 import _lib  # noqa: E402
from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch, synthetic_latent_batch  # noqa: E402

DEV = "cuda:0"


def timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) * 1e-3 / reps


def euler_rate(model, mb, steps):
    eng = model.engine()
    pb = eng.prepare(mb)
    grid = torch.linspace(0.0, 1.0, steps + 1)
    eng.rollout_fixed(pb, mb.x0.contiguous(), grid[:4], method="euler", save_frames=False)      # warm-up
    _, sec = timed(lambda: eng.rollout_fixed(pb, mb.x0.contiguous(), grid, method="euler", save_frames=False))
    eng.status()
    return pb.n_mol * steps / sec, sec / steps


def kernel_shares(fn):
    """Per-kernel-class device time (tib_profile_begin/end) of one call of fn."""
    import ctypes as C
    lib = _lib.load()
    ms = (C.c_double * _lib.N_KERNEL_KINDS)()
    n = (C.c_uint64 * _lib.N_KERNEL_KINDS)()
    lib.tib_profile_begin()
    fn()
    torch.cuda.synchronize()
    lib.tib_profile_end(ms, n)
    return {k: dict(ms=round(ms[i], 3), launches=int(n[i])) for i, k in enumerate(_lib.KERNEL_KINDS) if n[i]}


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="cfg3,dopri5,f256,div,divdopri5,adw")
    ap.add_argument("--f256-mols", type=int, default=512)
    only = set(ap.parse_args().only.split(","))
    if "cfg3" in only:
        bench_cfg3()
    if "dopri5" in only:
        bench_dopri5()
    if "f256" in only:
        bench_f256(ap.parse_args().f256_mols)
    if "div" in only:
        bench_div()
    if "divdopri5" in only:
        bench_div_dopri5()
    if "adw" in only:
        bench_adw()


def bench_cfg3():
    # ---- cfg 3: latent multi-T flow on mixed molecule sizes, F = 128 (fused tcgen05 kernels) and F = 256 (layered tcgen05 path)
    from thermodynamic_interpolation_b200.latent.models.cpainn import cPaiNN as Latent
    gen = torch.Generator().manual_seed(2)
    n_list = torch.randint(9, 26, (16384,), generator=gen).tolist()
    for F, n_mol, steps in ((128, 16384, 20), (256, 4096, 5)):
        torch.manual_seed(0)
        model = perturb_(Latent(n_features=F, score_layers=5, temp_length=75), 1).eval().to(DEV)
        mb = synthetic_latent_batch(n_mol, n_list[:n_mol], T=800, seed=3).to(DEV)
        rate, per = euler_rate(model, mb, steps)
        print(json.dumps(dict(workload=f"cfg 3: latent multi-T, {n_mol} molecules with 9..25 atoms, F={F} L=5, Euler",
                              math=_lib.MATH_NAMES[1], value=rate, unit="molecule*steps/s", ms_per_step=per * 1e3,
                              n_nodes=int(mb.x0.shape[0]), n_edges=int(mb.edge_index.shape[1]))), flush=True)
        del model, mb


def bench_dopri5():
    # ---- cfg 2 under dopri5
    from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN as Ambient
    torch.manual_seed(0)
    model = perturb_(Ambient(n_features=128, score_layers=5, temp_length=100), 1).eval().to(DEV)
    mb = synthetic_ambient_batch(4096, 9, seed=100).to(DEV)
    integ = MoleculeIntegrator(model, method="dopri5", n_step=100, atol=1e-5, rtol=1e-5)
    integ.rollout(mb)
    (xts, dlogp, nfe, _), sec = timed(lambda: integ.rollout(mb))
    print(json.dumps(dict(workload="cfg 2 under the reference's solver: dopri5 rtol=atol=1e-5, 100 frames, 4096 x 9 atoms, F=128",
                          math=_lib.MATH_NAMES[1], nfe=int(nfe), attempts=integ.last_stats["attempts"],
                          accepted=integ.last_stats["accepted"], seconds=sec, value=4096 * nfe / sec,
                          unit="molecule*drift-evals/s", finite=bool(torch.isfinite(xts).all()))), flush=True)


def bench_f256(n_mol):
    # ---- 10506-shaped: 25 atoms, F = 256 (config/ambient/10506_settings_no_900.json:14): layered tcgen05 path vs fp32 kernels
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN as Ambient
    torch.manual_seed(0)
    model = perturb_(Ambient(n_features=256, score_layers=5, temp_length=100), 1).eval().to(DEV)
    mb = synthetic_ambient_batch(n_mol, 25, seed=100).to(DEV)
    for math in (1, 0):
        model.set_math(math)
        rate, per = euler_rate(model, mb, 5)
        eng = model.engine()
        pb = eng.prepare(mb)
        shares = kernel_shares(lambda: eng.drift(pb, mb.x0.contiguous(), 0.3))
        flops = 25 * 12 * 256 ** 2 + 5 * (600 * 30 * 256 ** 2 + 25 * 24 * 256 ** 2) + 25 * (4 * 256 ** 2)
        print(json.dumps(dict(workload=f"10506-shaped: {n_mol} conformers x 25 atoms, F=256 L=5, Euler", math=_lib.MATH_NAMES[math],
                              value=rate, unit="molecule*steps/s", ms_per_step=per * 1e3, algorithmic_tflops=rate * flops / 1e12,
                              kernel_ms_per_drift=shares)), flush=True)


def bench_div():
    # ---- cfg 2 with the exact divergence (return_dlogp=True right-hand side)
    import time
    from oracle import cpainn_oracle as co   # CPU baseline leg only
    from tests._util import oracle_hp_sd, oracle_temps
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN as Ambient
    torch.manual_seed(0)
    model = perturb_(Ambient(n_features=128, score_layers=5, temp_length=100), 1).eval().to(DEV)
    mb = synthetic_ambient_batch(4096, 9, seed=100).to(DEV)
    eng = model.engine()
    pb = eng.prepare(mb)
    eng.drift_div(pb, mb.x0.contiguous(), 0.3)
    (b, div), sec = timed(lambda: eng.drift_div(pb, mb.x0.contiguous(), 0.3), reps=2)
    shares = kernel_shares(lambda: eng.drift_div(pb, mb.x0.contiguous(), 0.3))
    _, sec_drift = timed(lambda: eng.drift(pb, mb.x0.contiguous(), 0.3), reps=3)
    small = synthetic_ambient_batch(32, 9, seed=100)
    hp, sd = oracle_hp_sd(model)
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    t0 = time.perf_counter()
    dref = co.divergence(sd, hp, small.x0, 0.3, small.atoms, small.edge_index, small.edge_type, small.ptr.tolist(),
                         **oracle_temps(small, hp))
    cpu_sec = time.perf_counter() - t0
    err = float((div[:32].cpu() - dref).abs().max() / dref.abs().max())
    print(json.dumps(dict(workload="cfg 2 right-hand side with return_dlogp=True: drift + exact divergence, 4096 x 9 atoms, F=128 L=5",
                          math="tangent GEMMs on tcgen05 (split-f16 x3), layered MLP-chain kernels", seconds_per_eval=sec,
                          seconds_per_plain_drift=sec_drift, ratio_to_plain_drift=sec / sec_drift, kernel_ms=shares,
                          value=4096 / sec, unit="molecule*(drift+divergence) evals/s",
                          cpu_oracle=dict(value=32 / cpu_sec, seconds=cpu_sec, sample="32 conformers, autograd, "
                                          f"{torch.get_num_threads()} threads"),
                          max_rel_diff_vs_oracle_on_sample=err)), flush=True)


def bench_div_dopri5():
    # ---- the reference's production configuration: return_dlogp: 1 under dopri5, rtol = atol = 1e-5, 100 frames
    #      (config/ambient/00031_settings_no_300.json:29,34-36) at cfg-2 size
    from thermodynamic_interpolation_b200.ambient.integrators import MoleculeIntegrator
    from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN as Ambient
    torch.manual_seed(0)
    model = perturb_(Ambient(n_features=128, score_layers=5, temp_length=100), 1).eval().to(DEV)
    mb = synthetic_ambient_batch(4096, 9, seed=100).to(DEV)
    integ = MoleculeIntegrator(model, method="dopri5", n_step=100, atol=1e-5, rtol=1e-5, return_dlogp=True)
    (xts, dlogp, nfe, _), sec = timed(lambda: integ.rollout(mb))
    print(json.dumps(dict(workload="cfg 2 with return_dlogp=True under the reference's solver: dopri5 rtol=atol=1e-5, 100 frames, 4096 x 9 atoms, F=128 L=5",
                          math="tensor-core tangents, tuple-state dopri5 inside libtib.so", nfe=int(nfe),
                          attempts=integ.last_stats.get("attempts"), accepted=integ.last_stats.get("accepted"), seconds=sec,
                          value=4096 * nfe / sec, unit="molecule*(drift+divergence) evals/s",
                          finite=bool(torch.isfinite(xts).all() and torch.isfinite(dlogp).all()),
                          dlogp_mean=float(dlogp[-1].mean()), dlogp_std=float(dlogp[-1].std()))), flush=True)


def bench_adw():
    # ---- cfg 1: asymmetric double well (the reference's CPU configuration)
    import time
    from oracle import cpainn_oracle as co   # CPU baseline leg only
    from thermodynamic_interpolation_b200.adw.integrators import StandardIntegrator
    from thermodynamic_interpolation_b200.adw.models.simple import FCNetMultiBeta
    torch.manual_seed(5)
    adw = perturb_(FCNetMultiBeta(1, 1, 256, 5).double(), 6).eval()
    gen = torch.Generator().manual_seed(7)
    x0 = torch.randn(10000, 1, generator=gen)
    b0 = torch.full((10000, 1), 1.0, dtype=torch.float64)
    b1 = torch.full((10000, 1), 1.25, dtype=torch.float64)
    t0 = time.perf_counter()
    xo, dlo = co.adw_rollout(adw.state_dict(), x0, b0, b1, method="euler", n_step=101)
    cpu_sec = time.perf_counter() - t0
    adw = adw.to(DEV)
    integ = StandardIntegrator(adw, method="euler", n_step=101, return_dlogp=True)
    args = (x0.to(DEV), b0.to(DEV), b1.to(DEV))
    integ.rollout(*args)
    (x, dl), sec = timed(lambda: integ.rollout(*args), reps=3)
    err = float((x[-1].cpu().double() - xo[-1].double()).abs().max())
    print(json.dumps(dict(workload="cfg 1: ADW FCNetMultiBeta fp64, 10000 samples x 100 Euler steps with exact divergence",
                          seconds=sec, value=10000 * 100 / sec, unit="sample*steps/s",
                          cpu_oracle=dict(value=10000 * 100 / cpu_sec, seconds=cpu_sec, threads=torch.get_num_threads()),
                          max_abs_diff_final_x=err)), flush=True)


if __name__ == "__main__":
    main()
