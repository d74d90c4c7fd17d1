"""Build recipe of libharmonica_b200.so (nvcc, sm_100a only, in-tree)."""

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = [os.path.join(_HERE, "csrc", "hb200_api.cu")]
HEADERS = [
    os.path.join(_HERE, "csrc", "hb200_math.cuh"),
    os.path.join(_HERE, "csrc", "hb200_fast.cuh"),
    os.path.join(_HERE, "csrc", "hb200_xmath.cuh"),
    os.path.join(_HERE, "csrc", "hb200_tables.h"),
    os.path.join(_HERE, "csrc", "hb200_kernels.cuh"),
    os.path.join(_HERE, "csrc", "hb200_tess.cuh"),
    os.path.join(_HERE, "csrc", "hb200_tess_leaves.cuh"),
    os.path.join(_HERE, "csrc", "hb200_trig.cuh"),
    os.path.join(_HERE, "csrc", "hb200_fit_kernels.cuh"),
    os.path.join(_HERE, "csrc", "hb200_fit_host.cuh"),
    os.path.join(os.path.dirname(_HERE), "include", "harmonica_b200.h"),
]
OUTPUT = os.path.join(_HERE, "libharmonica_b200.so")

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
]  # fmt: skip


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(OUTPUT):
        return False
    built = os.path.getmtime(OUTPUT)
    return all(os.path.getmtime(p) <= built for p in SOURCES + HEADERS + [__file__])


def build_library(force=False, verbose=False):
    """Compile the CUDA library in-tree; returns the path of the .so."""
    if not force and up_to_date():
        return OUTPUT
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", OUTPUT, *SOURCES]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    env = dict(os.environ)
    env.pop("CC", None)  # the image's CC lacks some spec files; let nvcc pick the system g++
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return OUTPUT
