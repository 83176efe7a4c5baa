"""Per-launch timing of the training step's GEMMs (TIB_TRAIN_TRACE=1: synchronous CUDA-event timing inside libtib.so),
aggregated by shape:  TIB_TRAIN_TRACE=1 python tools/train_trace.py [n_mol]"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from thermodynamic_interpolation_b200.ambient.interpolants import LinearInterpolant
from thermodynamic_interpolation_b200.batch import synthetic_train_batches
from thermodynamic_interpolation_b200.synthetic import seeded_ambient_model
from thermodynamic_interpolation_b200.train_ambient import Trainer
n = int(sys.argv[1])
model = seeded_ambient_model(128, 5, 100, seed=0).to("cuda:0")
tr = Trainer(model, LinearInterpolant(a=1, gamma="sin2"))
b0, b1 = synthetic_train_batches(n, 9, seed=3)
for i in range(2):
    if i == 1:
        print("=== step", file=sys.stderr, flush=True)
    tr.step(b0, b1)
torch.cuda.synchronize()
''' % REPO


def main():
    n = sys.argv[1] if len(sys.argv) > 1 else "256"
    env = dict(os.environ, TIB_TRAIN_TRACE="1")
    err = subprocess.run([sys.executable, "-c", CHILD, n], env=env, capture_output=True, text=True).stderr
    err = err.split("=== step")[-1]
    agg = collections.OrderedDict()
    pat = re.compile(r"\[gemm\] (M=\d+ N=\d+ K=\d+ tA=\d tB=\d idxA=\d idxB=\d mode=\d splits=\d+ ctas=\d+)\s+([\d.]+) us\s+([\d.]+) TFLOP/s")
    tot = 0.0
    for m in pat.finditer(err):
        a = agg.setdefault(m.group(1), [0, 0.0, 0.0])
        a[0] += 1; a[1] += float(m.group(2)); a[2] = float(m.group(3))
        tot += float(m.group(2))
    for k, (cnt, us, tf) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / 1e3:8.3f} ms {100 * us / tot:5.1f}%  {cnt:3d} x {us / cnt:7.1f} us  {tf:6.1f} TFLOP/s  {k}")
    print(f"{tot / 1e3:8.3f} ms in GEMMs per step ({n} molecules)")
    agg2, tot2 = collections.OrderedDict(), 0.0
    for m in re.finditer(r"\[kern\] (\w+)\s+([\d.]+) us", err):
        a = agg2.setdefault(m.group(1), [0, 0.0])
        a[0] += 1; a[1] += float(m.group(2)); tot2 += float(m.group(2))
    for k, (cnt, us) in sorted(agg2.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / 1e3:8.3f} ms {100 * us / max(tot2, 1e-9):5.1f}%  {cnt:3d} x {us / cnt:7.1f} us  {k}")
    print(f"{tot2 / 1e3:8.3f} ms in the other kernels per step")


if __name__ == "__main__":
    main()
