"""Builds trueno_b200/libtrueno_cuda.so (the C-ABI library) with nvcc for sm_100a only.

In-tree build so the .so travels to the GPU box with the repo snapshot.  nvcc cross-compiles
without a GPU.  Usage: python trueno_b200/build.py [--force] [--verbose]   (run as a script: importing the package
first would need the library this script is about to build)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtrueno_cuda.so")
OBJ_DIR = os.path.join(HERE, "_build")
SOURCES = ["context.cu", "reduce.cu", "map.cu", "softmax.cu", "gemm_simt.cu", "gemm_tc.cu", "batch.cu", "peer.cu", "conv.cu", "attention.cu", "eigen.cu", "gather.cu", "api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-cudart", "shared",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tcgen05.cuh"), os.path.join(ROOT, "include", "trueno_cuda.h"), __file__]
    objs, procs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [NVCC, *FLAGS, "-c", path, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"[{src}]\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-cudart", "shared", "-o", LIB, *objs,
               "-Xlinker", "-rpath,/usr/local/cuda/lib64", "-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
