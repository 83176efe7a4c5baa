"""Times tib_gemm_f16x3 (csrc/train_gemm.cuh) on the shapes of the training step with CUDA events and prints the clock64()
timeline of CTA (0,0,0) (tib_gemm_debug):  python tools/gemm_probe.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thermodynamic_interpolation_b200 import _lib  # noqa: E402
from thermodynamic_interpolation_b200.train import gemm_f16x3  # noqa: E402

dev = "cuda:0"
lib = _lib.load()


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def timeline(fn):
    lib.tib_gemm_debug(1, None)
    fn()
    buf = (C.c_longlong * 64)()
    lib.tib_gemm_debug(1, buf)
    n = int(buf[63])
    t = [buf[i] - buf[0] for i in range(n)]
    lib.tib_gemm_debug(0, None)
    return t


def main():
    F = 128
    cases = []
    for rows in (4608, 18432, 36864):
        X = torch.randn(rows, 2 * F, device=dev)
        W5 = torch.randn(5 * F, 2 * F, device=dev) * 0.1
        dY = torch.randn(rows, 5 * F, device=dev) * 1e-3
        amax = dY.abs().max().reshape(1)
        out1 = torch.zeros(rows, F, device=dev)
        out5 = torch.zeros(rows, 5 * F, device=dev)
        g = torch.zeros(5 * F, F, device=dev)
        cases.append((f"fwd  [{rows}x128] = X[{rows}x128] W^T", 2.0 * rows * F * F,
                      lambda X=X, W5=W5, out1=out1: gemm_f16x3(X[:, :F], W5[:F, :F], out=out1)))
        cases.append((f"fwd  [{rows}x640] = H[{rows}x128] W3^T", 2.0 * rows * 5 * F * F,
                      lambda X=X, W5=W5, out5=out5: gemm_f16x3(X[:, :F], W5[:, :F], out=out5)))
        cases.append((f"dgrad[{rows}x128] = dY[{rows}x640] W3", 2.0 * rows * 5 * F * F,
                      lambda dY=dY, W5=W5, out1=out1, amax=amax: gemm_f16x3(dY, W5[:, :F], trans_b=True, amax_a=amax, out=out1)))
        cases.append((f"wgrad[640x128] = dY^T X over {rows} rows", 2.0 * rows * 5 * F * F,
                      lambda dY=dY, X=X, g=g, amax=amax, rows=rows: gemm_f16x3(dY, X[:, :F], trans_a=True, trans_b=True, amax_a=amax, out=g,
                                                                  mode=_lib.GEMM_ATOMIC, split_k=True, M=5 * F, N=F, K=rows)))
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, flops, fn in cases:
        if only and only not in name:
            continue
        print(f"{name:48s}", end=" ", flush=True)
        us = timeit(fn)
        print(f"{us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s", end=" ", flush=True)
        print(f"timeline(cycles) {timeline(fn)}", flush=True)


if __name__ == "__main__":
    main()
