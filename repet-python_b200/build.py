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
# compiled once per window length (REPET_WIN_N = 512 / 1024 / 2048: sampling rates up to 51.2 kHz)
PER_WINDOW_SOURCES = ["repet_kernels.cu", "repet_sim.cu", "repet_simgemm.cu", "repet_drivers.cu", "repet_helpers.cu"]
WINDOW_LENGTHS = [512, 1024, 2048]
# compiled once: handle lifetime and the extern "C" dispatch
COMMON_SOURCES = ["repet_abi.cu", "repet_generic.cu"]
SOURCES = PER_WINDOW_SOURCES + COMMON_SOURCES
HEADERS = ["repet_kernels.cuh", "fft_core.cuh", "median_networks.cuh", "median_networks_large.cuh", "repet_internal.h", os.path.join("..", "..", "include", "repet_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]
OBJ_DIR = os.path.join(HERE, "build")


def _stale():
    if not os.path.isfile(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps)


def _compile(job):
    src, win_n, obj, verbose = job
    cmd = [NVCC] + COMPILE_FLAGS + (["-Xptxas", "-v"] if verbose else []) + os.environ.get("REPET_EXTRA_NVCC_FLAGS", "").split()
    if win_n:
        cmd += ["-DREPET_WIN_N=%d" % win_n]
    cmd += ["-c", os.path.join(CSRC, src), "-o", obj]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return proc.returncode, " ".join(cmd), proc.stdout


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = []
    for n in WINDOW_LENGTHS:
        for src in PER_WINDOW_SOURCES:
            jobs.append((src, n, os.path.join(OBJ_DIR, "%s.w%d.o" % (src[:-3], n)), verbose))
    for src in COMMON_SOURCES:
        jobs.append((src, 0, os.path.join(OBJ_DIR, "%s.o" % src[:-3]), verbose))
    workers = max(1, min(len(jobs), (os.cpu_count() or 4)))
    with ThreadPoolExecutor(workers) as pool:
        results = list(pool.map(_compile, jobs))
    log = []
    for rc, cmd, out in results:
        if rc != 0:
            raise RuntimeError("nvcc failed:\n" + cmd + "\n" + out)
        log.append(out)
    cmd = [NVCC] + LINK_FLAGS + [j[2] for j in jobs] + ["-o", LIB]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
    if verbose:
        print("".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
