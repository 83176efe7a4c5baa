"""Builds libtib.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtib.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                 # FMAs are explicit fmaf(); see csrc/common.cuh
    "-shared", "-Xcompiler", "-fPIC",
]


def sources():
    """The library's translation units: the sampling path and the training step."""
    return [os.path.join(CSRC, "tib_api.cu"), os.path.join(CSRC, "train_api.cu")]


OBJ_DIR = os.path.join(HERE, "build")

# headers each translation unit includes (an object is rebuilt when one of them is newer)
TRAIN_HEADERS = ("train.cuh", "train_gemm.cuh", "tc_common.cuh", "common.cuh")


def _unit_deps(src):
    name = os.path.basename(src)
    hdr = [os.path.join(HERE, "..", "include", "tib.h")]
    if name == "train_api.cu":
        return [src] + [os.path.join(CSRC, h) for h in TRAIN_HEADERS] + hdr
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)
            if f not in ("train_api.cu", "train.cuh", "train_gemm.cuh")] + hdr


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(HERE, "..", "include", "tib.h"))
    return out


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libtib.so cannot be built")
    os.makedirs(OBJ_DIR, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and all(os.path.getmtime(p) <= os.path.getmtime(obj) for p in _unit_deps(src)):
            continue
        cmd = [nvcc, *compile_flags, "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for cmd, proc in jobs:          # the translation units compile side by side
        out, err = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out + err)
        if verbose:
            print(err, file=sys.stderr)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB + ".tmp", *objs]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
