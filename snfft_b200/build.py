"""Build the CUDA library in-tree: ``python -m snfft_b200.build``.

nvcc cross-compiles for sm_100a without a GPU; the resulting
``snfft_b200/_lib/libsnfft_b200.so`` is git-ignored but travels with the working tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
OUT = os.path.join(OUT_DIR, "libsnfft_b200.so")
SOURCES = ["plan.cpp", "kernels.cu", "capi.cu"]
HEADERS = ["plan.hpp", "device_plan.hpp", "kernels.cuh", "gen_programs.cpp",
           os.path.join("..", "..", "include", "snfft_b200.h")]
GENERATED = os.path.join(CSRC, "gen_small.cuh")


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return all(os.path.getmtime(d) <= t for d in deps)


def generate() -> str:
    """Regenerate csrc/gen_small.cuh (straight-line register programs of levels 2..6) from the
    plan builder: g++ gen_programs.cpp plan.cpp && ./gen_programs gen_small.cuh."""
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "gen_programs")
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(CSRC, "gen_programs.cpp"),
                        os.path.join(CSRC, "plan.cpp")], check=True)
        fresh = os.path.join(tmp, "gen_small.cuh")
        subprocess.run([exe, fresh], check=True)
        new = open(fresh).read()
    if not os.path.exists(GENERATED) or open(GENERATED).read() != new:
        with open(GENERATED, "w") as fh:
            fh.write(new)
    return GENERATED


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    generate()
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
           "-lineinfo", "-Xcompiler", "-fPIC", "-shared", "-o", OUT]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
