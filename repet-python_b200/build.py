"""Build librepet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python repet-python_b200/build.py [--force]

The library is self-contained (static cudart): it is loaded by ctypes from
`repet-python_b200/repet/_host.py`, by `bench.py` and by the tests, and travels to the GPU
box with the repository snapshot.
"""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "librepet_b200.so")
SOURCES = ["repet_kernels.cu", "repet_sim.cu", "repet_simgemm.cu", "repet_abi.cu", "repet_drivers.cu"]
HEADERS = ["repet_kernels.cuh", "fft2048.cuh", "median_networks.cuh", "median_networks_large.cuh", "repet_internal.h", os.path.join("..", "..", "include", "repet_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def _stale():
    if not os.path.isfile(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, f) for f in SOURCES] + ["-o", LIB]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
    if verbose:
        print(proc.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
