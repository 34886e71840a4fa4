"""Build fedoo_b200/_fdk.so with nvcc for sm_100a (in-tree, so it travels with the repo snapshot)."""

from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "fdk_api.cu")
OUT = os.path.join(HERE, "_fdk.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [
    os.path.join(ROOT, "include", "fdk.h")
]


def nvcc_command(out=OUT):
    return [
        os.environ.get("NVCC", "nvcc"),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
        "-Xcompiler", "-fPIC", "-shared",
        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc"),
        "-o", out, SRC,
    ]  # fmt: skip


def build(force=False, verbose=True):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
