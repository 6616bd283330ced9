"""
Golden vectors at the sizes of BASELINE.json configs 3-5, from the UNMODIFIED reference.

TEST INFRASTRUCTURE; run in the build container only (needs /root/reference; minutes to an hour of
CPU per case, several GB of memory):

    python oracle/make_golden_long.py --case adaptive_10min     # ~2 min
    python oracle/make_golden_long.py --case extended_1h        # ~6 min
    python oracle/make_golden_long.py --case sim_5min           # ~10 min
    python oracle/make_golden_long.py --case sim_10min          # ~40 min, ~12 GB
    python oracle/make_golden_long.py --case simonline_1h       # ~60 min, ~25 GB

Every case runs the reference driver `repet.<fn>` (through `oracle/reference_shim.py`) on a seeded synthetic
track with its integer decisions recorded (the helpers `_periods`, `_indices`, `_localmaxima` are wrapped, not
changed), cross-checks the oracle (`oracle/repet_oracle.py`) against it -- signal to 1e-12 of the peak (in practice
bit-identical), integers exactly -- and stores in `tests/golden/long_<case>.npz` the REFERENCE's outputs:
  * the integer outputs in full or as digests -- per-frame periods (adaptive), per-segment periods
    (extended), per-frame list lengths + one digest per block of 64 lists + the total (sim, simonline);
  * rms, peak and every DECIMATE-th sample of the background signal.
The inputs are regenerated from their seeds at test time (`long_input` below, shared with tests/).
"""

import argparse
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "repet-python_b200"))

import repet_oracle as oracle  # noqa: E402
import repet_synth  # noqa: E402
from make_golden import list_digests  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
FS = 44100
DECIMATE = 4099
BLOCK = 64

# name -> driver, samples, seed of the synthetic track, dither (0 = the float64 samples are fp32-representable)
CASES = {
    # BASELINE configs[2]: repet.adaptive on a 10-minute stereo track (T = 25 841 frames)
    "adaptive_10min": dict(fn="adaptive", samples=600 * FS, index=7000, dither=0.0),
    # BASELINE configs[4]: repet.extended on one hour (719 segments)
    "extended_1h": dict(fn="extended", samples=3600 * FS, index=7001, dither=0.0),
    # BASELINE configs[3]: repet.sim on 5- and 10-minute tracks (T = 12 921 / 25 841); float64 samples that are
    # NOT fp32-representable (1e-9 dither), as a user's array would be
    "sim_5min": dict(fn="sim", samples=300 * FS, index=4243, dither=1e-9),
    "sim_10min": dict(fn="sim", samples=600 * FS, index=4244, dither=1e-9),
    # BASELINE configs[4]: repet.simonline on one hour; S = 155 038 * 1024 + 2048 so that the reference's
    # accidental 2-D pad is empty and it can run at all
    "simonline_1h": dict(fn="simonline", samples=155038 * 1024 + 2048, index=7002, dither=0.0),
    # small versions of the same cases for quick checks of this script
    "adaptive_1min": dict(fn="adaptive", samples=60 * FS, index=7000, dither=0.0),
    "simonline_1min": dict(fn="simonline", samples=2582 * 1024 + 2048, index=7002, dither=0.0),
}


def long_input(case):
    """(S, C) float64 input of a long case.  Shared with tests/ (imported from there)."""
    spec = CASES[case]
    x = repet_synth.make_clip(spec["index"], spec["samples"], redraw_seconds=(60, 120)).T.astype(np.float64)
    if spec["dither"]:
        rng = np.random.default_rng(99 + spec["index"])
        x += spec["dither"] * rng.standard_normal(x.shape)
    return x


def run_reference(ref, fn, x):
    """The reference driver with its integer decisions recorded: `_periods` (original per segment / adaptive per
    frame), `_indices` (sim) and `_localmaxima` (simonline, slot indices mapped to frame indices exactly as the
    driver's ring buffer holds them, repet.py:837-866) are wrapped, not changed."""
    captured = {"periods": [], "lists": []}
    saved = {name: getattr(ref, name) for name in ("_periods", "_indices", "_localmaxima")}
    N, w, H = oracle.stft_parameters(FS)
    buffer_frames = round((ref.buffer_length * FS) / H)

    def periods(*a, **k):
        r = saved["_periods"](*a, **k)
        captured["periods"].append(np.asarray(r).copy())
        return r

    def indices(*a, **k):
        r = saved["_indices"](*a, **k)
        captured["lists"] = [np.asarray(v).copy() for v in r]
        return r

    def localmaxima(*a, **k):
        values, idx = saved["_localmaxima"](*a, **k)
        j = buffer_frames - 1 + len(captured["lists"])
        j0 = j % buffer_frames
        slots = np.asarray(idx)
        captured["lists"].append(np.where(slots <= j0, j - (j0 - slots), j - (j0 - slots) - buffer_frames))
        return values, idx

    ref._periods = periods
    if fn == "sim":
        ref._indices = indices
    if fn == "simonline":
        ref._localmaxima = localmaxima
    try:
        y = getattr(ref, fn)(x, FS)
    finally:
        for name, f in saved.items():
            setattr(ref, name, f)
    return y, captured


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True, choices=sorted(CASES))
    ap.add_argument("--no-oracle", action="store_true", help="skip the oracle cross-check (simonline_1h: the "
                    "frame-parallel oracle needs ~25 GB and an hour; the 1-minute case checks the same code)")
    args = ap.parse_args()
    warnings.simplefilter("ignore")
    import reference_shim

    ref = reference_shim.load()
    assert ref is not None, "reference not present at /root/reference"
    spec = CASES[args.case]
    fn = spec["fn"]
    t0 = time.perf_counter()
    x = long_input(args.case)
    print("%s: input %s in %.0f s" % (args.case, x.shape, time.perf_counter() - t0), flush=True)

    t0 = time.perf_counter()
    y_ref, captured = run_reference(ref, fn, x)
    print("%s: reference driver in %.0f s" % (args.case, time.perf_counter() - t0), flush=True)
    out = {}
    if fn == "adaptive":
        assert len(captured["periods"]) == 1
        out["periods"] = captured["periods"][0].astype(np.int16)
    elif fn == "extended":
        out["periods"] = np.array([int(p) for p in captured["periods"]], dtype=np.int16)
    else:
        lists = captured["lists"]
        out["counts"] = np.array([len(v) for v in lists], dtype=np.uint8)
        out["digests"] = list_digests(lists, BLOCK)
        out["total"] = np.int64(sum(len(v) for v in lists))
        if fn == "simonline":
            N, w, H = oracle.stft_parameters(FS)
            out["first_frame"] = np.int64(round((ref.buffer_length * FS) / H) - 1)

    if not args.no_oracle:
        t0 = time.perf_counter()
        y_orc, det = getattr(oracle, fn)(x, FS, return_details=True)
        print("%s: oracle in %.0f s" % (args.case, time.perf_counter() - t0), flush=True)
        peak = float(np.max(np.abs(y_ref)))
        err = float(np.max(np.abs(y_ref - y_orc))) / peak
        print("%s: max |reference - oracle| / peak = %.3e" % (args.case, err), flush=True)
        assert err <= 1e-12, err
        if fn in ("adaptive", "extended"):
            assert np.array_equal(np.asarray(det["periods"]).astype(np.int16), out["periods"]), "periods differ"
        else:
            assert len(det["indices"]) == len(lists)
            assert all(np.array_equal(a, b) for a, b in zip(det["indices"], lists)), "oracle lists differ from the reference's"
        out["oracle_checked"] = np.array(True)
    else:
        out["oracle_checked"] = np.array(False)
    out["rms"] = np.sqrt(np.mean(np.square(y_ref)))
    out["max"] = np.max(np.abs(y_ref))
    out["dec"] = y_ref[::DECIMATE].copy()
    out["decimate"] = np.int64(DECIMATE)
    out["samples"] = np.int64(x.shape[0])
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, "long_%s.npz" % args.case)
    np.savez_compressed(path, **out)
    print("%s: written %s (%d bytes)" % (args.case, path, os.path.getsize(path)), flush=True)


if __name__ == "__main__":
    main()
