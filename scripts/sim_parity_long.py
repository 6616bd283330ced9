#!/usr/bin/env python
"""Long-track parity probe for REPET-SIM's similar-frame lists.

On the GPU box:   python scripts/sim_parity_long.py gpu  [seconds]   -> gpurun_out/sim_long_<seconds>s.npz
In the build container (no GPU; /root/reference present):
                  python scripts/sim_parity_long.py check [seconds]  -> compares with the reference's lists

`gpu` stores the lists of `repet.sim` on a seeded synthetic track with the similarity operand taken
from the float64 front end (k_frames64, the default) and from the fp32 magnitudes of k_stft
(sim_frames64 = 0).  `check` runs the unmodified reference (oracle/reference_shim.py) on the same
track -- minutes of pure-Python `_localmaxima` -- and counts the lists that differ.  The float64
input deliberately is NOT representable in fp32 (a 1e-9 dither is added), as a user's float64
array would be.
"""

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "repet-python_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
FS = 44100


def track(seconds):
    import make_golden
    import repet_synth

    if int(seconds) == make_golden.SIM_LONG["seconds"]:
        return make_golden.sim_long_input()  # the track of tests/golden/sim_long.npz
    x = repet_synth.make_clip(4242, int(seconds * FS)).T.astype(np.float64)
    rng = np.random.default_rng(99)
    return x + 1e-9 * rng.standard_normal(x.shape)


def flatten(lists):
    counts = np.array([len(v) for v in lists], dtype=np.int32)
    flat = np.concatenate(lists).astype(np.int32) if len(lists) else np.zeros(0, np.int32)
    return counts, flat


def differing(counts_a, flat_a, counts_b, flat_b):
    offs_a = np.concatenate([[0], np.cumsum(counts_a)])
    offs_b = np.concatenate([[0], np.cumsum(counts_b)])
    bad = []
    for i in range(len(counts_a)):
        if counts_a[i] != counts_b[i] or not np.array_equal(flat_a[offs_a[i] : offs_a[i + 1]], flat_b[offs_b[i] : offs_b[i + 1]]):
            bad.append(i)
    return bad


def report(path, counts, flat):
    if not os.path.isfile(path):
        print("no GPU lists at", path)
        return
    got = np.load(path)
    for name in ("f64", "f32"):
        bad = differing(got[name + "_counts"], got[name + "_flat"], counts, flat)
        print("%s front end vs reference: %d of %d lists differ %s" % (name, len(bad), len(counts), bad[:10]))


def main():
    mode = sys.argv[1]
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
    path = os.path.join(ROOT, "gpurun_out", "sim_long_%ds.npz" % int(seconds))
    x = track(seconds)
    if mode == "gpu":
        import repet

        out = {}
        for name, knob in (("f64", 1), ("f32", 0)):
            repet._host.set_tuning(sim_frames64=knob)
            t0 = time.perf_counter()
            y, lists = repet._host.sim_f64(x, FS, repet._tunables(), return_indices=True)
            out[name + "_counts"], out[name + "_flat"] = flatten(lists)
            out[name + "_rms"] = float(np.sqrt(np.mean(y * y)))
            print("%s front end: %d frames, %d indices, %.2f s" % (name, len(lists), len(out[name + "_flat"]), time.perf_counter() - t0))
        repet._host.set_tuning(sim_frames64=1)
        bad = differing(out["f64_counts"], out["f64_flat"], out["f32_counts"], out["f32_flat"])
        print("lists that differ between the two front ends: %d of %d" % (len(bad), len(out["f64_counts"])))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez_compressed(path, **out)
    else:
        import reference_shim

        ref = reference_shim.load()
        t0 = time.perf_counter()
        cache = os.path.join(ROOT, "gpurun_out", "sim_long_ref_%ds.npz" % int(seconds))
        if os.path.isfile(cache):
            counts, flat = np.load(cache)["counts"], np.load(cache)["flat"]
            report(path, counts, flat)
            return
        N = 2048
        import scipy.signal.windows

        w = scipy.signal.windows.hamming(N, sym=False)
        spec = np.mean(np.stack([np.abs(ref._stft(x[:, c], w, N // 2)[: N // 2 + 1]) for c in range(x.shape[1])], axis=2), axis=2)
        S = ref._selfsimilaritymatrix(spec)
        lists = ref._indices(S, ref.similarity_threshold, int(round(ref.similarity_distance * FS / (N // 2))), ref.similarity_number)
        counts, flat = flatten(lists)
        print("reference: %d frames, %d indices, %.1f s" % (len(lists), len(flat), time.perf_counter() - t0))
        np.savez_compressed(cache, counts=counts, flat=flat)
        report(path, counts, flat)


if __name__ == "__main__":
    main()
