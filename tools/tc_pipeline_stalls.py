"""Diagnostics: where the tensor-core message kernel's pipeline waits (per-CTA stall-cycle counters,
include/tib.h tib_debug_counters).  Run on a B200:  python tools/tc_pipeline_stalls.py [--math 1]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._util import perturb_  # noqa: E402
from thermodynamic_interpolation_b200 import _lib  # noqa: E402
from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN  # noqa: E402
from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--math", type=int, default=1)
ap.add_argument("--mols", type=int, default=4096)
args = ap.parse_args()
torch.manual_seed(0)
model = perturb_(cPaiNN(n_features=128, score_layers=5, temp_length=100), 1).eval().to("cuda:0").set_math(args.math)
mb = synthetic_ambient_batch(args.mols, 9, seed=2).to("cuda:0")
eng = model.engine()
pb = eng.prepare(mb)
lib = _lib.load()
for _ in range(3):
    eng.drift(pb, mb.x0, 0.3)
lib.tib_debug_counters(eng.handle, 1, None, 0)
eng.drift(pb, mb.x0, 0.3)
eng.status()
n = 148
buf_all = np.zeros((2048, 8), dtype=np.int64)
lib.tib_debug_counters(eng.handle, 1, buf_all.ctypes.data_as(C.c_void_p), 2048)
buf = buf_all[:n]
names = ["mma:wait weights", "mma:wait operands", "producer:wait free slot", "mma:wait acc drain", "mma:total",
         "mma:issuing tcgen05.mma", "mma:issuing tcgen05.commit", "epi(thread 0):wait accumulators"]
tot = buf[:, 4].mean()
print("last message launch (layer 5), mean cycles per CTA over", n, "CTAs; tiles per CTA ~", 4096 * 9 / 16 / n)
for i, nm in enumerate(names):
    print(f"  {nm:28s} {buf[:, i].mean():12.0f}  ({100 * buf[:, i].mean() / tot:5.1f}% of mma total)")

ph = buf_all[n:3 * n].reshape(n, 2, 8).mean(0)
tiles = 4096 * 9 / 16 / n
print("epilogue phases, mean cycles per tile (every epilogue thread runs the same sequence; columns: thread 0 | thread 256):")
for i, nm in enumerate(["tile tables+barrier", "E1 PE + E2 s[src]", "E3 w hid1 + E4 e rows", "E5 w hid2 + E6 phi hid1", "E7 phi hid2",
                        "output layer", "write-back"]):
    print(f"  {nm:28s} {ph[0, i] / tiles:10.0f} | {ph[1, i] / tiles:10.0f}")
print("  total per tile               %10.0f | %10.0f" % (ph[0].sum() / tiles, ph[1].sum() / tiles))

up = buf_all[1024:1024 + 2 * n].reshape(n, 16)
utiles = np.where(np.arange(n) < 288 - n, 2, 1)[:, None]
names_u = ["build planes 0,1", "wait vv,uv (x3)", "q2 += vv^2, next builds (x3)", "-", "q image", "wait layer 1", "LayerNorm 1",
           "wait layer 2", "LayerNorm 2", "wait g", "v update", "wait a,c", "-", "-", "s update + barrier"]
print("update kernel (last launch), thread 0, mean cycles per 128-node tile:")
for i, nm in enumerate(names_u):
    print(f"  {nm:28s} {(up[:, i] / utiles[:, 0]).mean():10.0f}")
print("  total per tile               %10.0f" % (up[:, :15].sum(1) / utiles[:, 0]).mean())
