"""
Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

TEST INFRASTRUCTURE; run in the build container only (needs /root/reference):

    python oracle/make_golden.py

For every case it (1) runs the reference `repet.py` through `oracle/reference_shim.py`,
(2) asserts that `oracle/repet_oracle.py` reproduces it (integers bit-exact, floats to
1e-12 relative -- in practice bit-identical), and (3) stores the REFERENCE's outputs:
integer outputs in full, float signals as rms + every DECIMATE-th sample, small helper
outputs in full.  Inputs are either regenerated from seeds at test time
(`repet_synth.make_clip`, `numpy.random.default_rng`) or, for the one real-audio vector,
stored as the int16 PCM of the reference's bundled `audio_file.wav`
(tests/golden/audio_file_int16.npz; BASELINE config 1's input).
"""

import hashlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "repet-python_b200"))

import reference_shim  # noqa: E402
import repet_oracle as oracle  # noqa: E402
import repet_synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
DECIMATE = 251
FS = 44100

# driver-level cases: name -> (input spec, functions)
DRIVER_CASES = {
    "wav_full": dict(kind="wav", start=0, stop=None, functions=["original", "extended", "adaptive", "sim", "simonline"]),
    "wav_5s": dict(kind="wav", start=5 * FS, stop=10 * FS, functions=["original", "extended", "adaptive", "sim"]),
    "synth_12s": dict(kind="synth", index=3, samples=12 * FS + 321, channels=2,
                      functions=["original", "extended", "adaptive", "sim", "simonline"]),
    "synth_21s": dict(kind="synth", index=5, samples=21 * FS, channels=2, functions=["original", "extended"]),
    "synth_mono_8s": dict(kind="synth", index=7, samples=8 * FS + 1000, channels=1,
                          functions=["original", "adaptive", "sim"]),
    "synth_30s": dict(kind="synth", index=11, samples=30 * FS, channels=2, functions=["original", "adaptive"]),
}


# other sampling rates: window lengths 512 (8 kHz), 1024 (16, 22.05 kHz) and 2048 with another
# hop / period mapping (48 kHz); repet.py:130 (quirk Q16).  Written to drivers_rates.npz.
RATE_CASES = {
    "synth16k_12s": dict(kind="synth", fs=16000, index=21, samples=12 * 16000 + 123, channels=2,
                         functions=["original", "extended", "adaptive", "sim", "simonline"]),
    "synth8k_14s": dict(kind="synth", fs=8000, index=23, samples=14 * 8000 + 57, channels=2,
                        functions=["original", "extended", "adaptive", "sim", "simonline"]),
    "synth22k_mono_9s": dict(kind="synth", fs=22050, index=25, samples=9 * 22050, channels=1,
                             functions=["original", "adaptive", "sim"]),
    "synth48k_6s": dict(kind="synth", fs=48000, index=27, samples=6 * 48000 + 5, channels=2,
                        functions=["original", "sim"]),
}


# inputs outside the fast kernels' shapes, served by the general float64 path (repet_general.cu): window lengths 4096
# and 8192 (96 / 192 kHz, repet.py:130), more than two channels (repet.py:152-155 loops over any number), and a period
# range beyond 1024 frames.  Written to drivers_general.npz.
GENERAL_CASES = {
    "synth96k_13s": dict(kind="synth", fs=96000, index=41, samples=13 * 96000 + 11, channels=2,
                         functions=["original", "adaptive", "sim", "simonline"]),
    "synth96k_17s": dict(kind="synth", fs=96000, index=42, samples=17 * 96000, channels=2, functions=["extended"]),
    "synth192k_5s": dict(kind="synth", fs=192000, index=43, samples=5 * 192000 + 3, channels=1, functions=["original", "sim"]),
    "synth44k_3ch_12s": dict(kind="synth", index=45, samples=12 * FS + 100, channels=3,
                             functions=["original", "extended", "adaptive", "sim", "simonline"]),
    "synth44k_5ch_8s": dict(kind="synth", index=47, samples=8 * FS + 9, channels=5, functions=["original", "adaptive", "sim"]),
    "synth16k_4ch_13s": dict(kind="synth", fs=16000, index=49, samples=13 * 16000, channels=4, functions=["original", "simonline"]),
}
# period range above 1024 frames (period_range[1] = 30 s -> 1292 frames at 44.1 kHz) on a 100 s clip
LONG_PERIOD = dict(index=51, samples=100 * FS, channels=2, period_range=[1, 30])


# a 2-minute track for REPET-SIM (T = 5169 frames, ~394k list entries): long enough that some pairs of
# similarities are tied to ~1e-8 -- an fp32 analysis front end flips list 3619 -- and the float64 input is
# deliberately NOT representable in fp32 (1e-9 dither), as a user's array would be.  Stored as per-list
# counts, one digest per block of 64 lists, and list 3619 in full (sim_long.npz).
SIM_LONG = dict(index=4242, seconds=120, dither_seed=99, dither=1e-9, block=64, hard_frame=3619)


def sim_long_input():
    x = repet_synth.make_clip(SIM_LONG["index"], SIM_LONG["seconds"] * FS).T.astype(np.float64)
    rng = np.random.default_rng(SIM_LONG["dither_seed"])
    return x + SIM_LONG["dither"] * rng.standard_normal(x.shape)


def list_digests(lists, block):
    """uint64 digest of every block of `block` consecutive lists (lengths and contents)."""
    out = []
    for lo in range(0, len(lists), block):
        h = hashlib.sha256()
        for v in lists[lo : lo + block]:
            h.update(np.asarray([len(v)], dtype="<i4").tobytes())
            h.update(np.asarray(v, dtype="<i4").tobytes())
        out.append(int.from_bytes(h.digest()[:8], "little"))
    return np.array(out, dtype=np.uint64)


def pin_sim_long(ref):
    x = sim_long_input()
    N, w, H = oracle.stft_parameters(FS)
    spec_ref = np.stack([np.abs(ref._stft(x[:, c], w, H)[0 : N // 2 + 1]) for c in range(x.shape[1])], axis=2)
    lists = ref._indices(ref._selfsimilaritymatrix(np.mean(spec_ref, axis=2)), 0, int(round(FS / H)), 100)
    out = {
        "counts": np.array([len(v) for v in lists], dtype=np.uint8),
        "digests": list_digests(lists, SIM_LONG["block"]),
        "hard_list": np.asarray(lists[SIM_LONG["hard_frame"]], dtype=np.int32),
        "total": np.int64(sum(len(v) for v in lists)),
    }
    np.savez_compressed(os.path.join(GOLDEN, "sim_long.npz"), **out)
    print("sim_long: %d lists, %d entries pinned" % (len(lists), int(out["total"])))


def case_fs(spec):
    return spec.get("fs", FS)


def case_input(spec, wav=None):
    """(S, C) float64 input of a driver case.  Shared with tests/ (imported from there)."""
    if spec["kind"] == "wav":
        pcm = wav[spec["start"] : spec["stop"]]
        return pcm / pow(2, pcm.itemsize * 8 - 1)  # repet.py:929
    fs = case_fs(spec)
    step_length = oracle.stft_parameters(fs)[2]
    clip = repet_synth.make_clip(spec["index"], spec["samples"], spec["channels"], fs, step_length)
    return clip.T.astype(np.float64)


def helper_inputs():
    """Small seeded inputs of the helper-level cases.  Shared with tests/."""
    rng = np.random.default_rng(20260117)
    d = {}
    d["signal"] = rng.standard_normal(5000)
    d["window"] = np.hamming(257)[:256].copy()  # any window works at helper level
    d["step"] = 128
    d["spectrogram"] = np.abs(rng.standard_normal((33, 300))) + 0.01
    # give it some periodicity so argmax is meaningful
    d["spectrogram"] += 0.8 * np.abs(np.sin(np.arange(300) * 2 * np.pi / 23.0))[None, :]
    d["vector"] = rng.standard_normal(400)
    d["vector"][50:55] = 3.0  # plateau: no strict maximum there (quirk Q7)
    d["periods_per_frame"] = rng.integers(5, 40, size=300)
    return d


def _close(a, b, what):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.kind in "iu" or b.dtype.kind in "iu":
        assert np.array_equal(a, b), what
        return
    scale = max(np.max(np.abs(a)) if a.size else 0.0, 1e-300)
    err = np.max(np.abs(a - b)) / scale if a.size else 0.0
    assert err <= 1e-12, (what, err)


def pin_drivers(ref, cases, wav=None):
    """Run the reference and the oracle on every case, assert they agree, return the reference's outputs."""
    drivers = {}
    for case, spec in cases.items():
        x = case_input(spec, wav)
        FS = case_fs(spec)
        for fn in spec["functions"]:
            key = "%s/%s" % (case, fn)
            y_ref = getattr(ref, fn)(x, FS)
            y_orc, det = getattr(oracle, fn)(x, FS, return_details=True)
            _close(y_ref, y_orc, key)
            drivers[key + "/rms"] = np.sqrt(np.mean(np.square(y_ref)))
            drivers[key + "/max"] = np.max(np.abs(y_ref))
            drivers[key + "/dec"] = y_ref[::DECIMATE].copy()
            # integer outputs: recompute from the REFERENCE's own helpers
            N, w, H = oracle.stft_parameters(FS)
            C = x.shape[1]
            if fn in ("original", "adaptive", "sim"):
                spec_ref = np.stack([np.abs(ref._stft(x[:, c], w, H)[0 : N // 2 + 1]) for c in range(C)], axis=2)
                pr2 = np.round(np.array(ref.period_range) * FS / H).astype(int)
            if fn == "original":
                p = ref._periods(ref._beatspectrum(np.power(np.mean(spec_ref, axis=2), 2)), pr2)
                assert int(p) == det["period"], key
                drivers[key + "/period"] = np.int64(p)
            elif fn == "adaptive":
                B = ref._beatspectrogram(np.power(np.mean(spec_ref, axis=2), 2), int(round(10 * FS / H)), int(round(5 * FS / H)))
                p = ref._periods(B, pr2)
                assert np.array_equal(p, det["periods"]), key
                drivers[key + "/periods"] = p.astype(np.int64)
            elif fn == "sim":
                Sm = ref._selfsimilaritymatrix(np.mean(spec_ref, axis=2))
                lists = ref._indices(Sm, 0, int(round(FS / H)), 100)
                assert all(np.array_equal(a, b) for a, b in zip(lists, det["indices"])), key
                drivers[key + "/index_counts"] = np.array([len(v) for v in lists], dtype=np.int64)
                drivers[key + "/index_flat"] = np.concatenate(lists).astype(np.int64)
            elif fn == "extended":
                drivers[key + "/periods"] = np.array(det["periods"], dtype=np.int64)
            elif fn == "simonline":
                drivers[key + "/index_counts"] = np.array([len(v) for v in det["indices"]], dtype=np.int64)
                drivers[key + "/index_flat"] = np.concatenate(det["indices"]).astype(np.int64)
                drivers[key + "/first_frame"] = np.int64(det["first_frame"])
            print("%-28s rms %.17g  pinned" % (key, drivers[key + "/rms"]))
    return drivers


def pin_general(ref, provenance):
    general = pin_drivers(ref, GENERAL_CASES)
    # the long period range: a module tunable of the reference, restored afterwards
    spec = LONG_PERIOD
    x = repet_synth.make_clip(spec["index"], spec["samples"], spec["channels"]).T.astype(np.float64)
    saved = ref.period_range
    ref.period_range = list(spec["period_range"])
    try:
        y_ref = ref.original(x, FS)
    finally:
        ref.period_range = saved
    y_orc, det = oracle.original(x, FS, return_details=True, period_range=tuple(spec["period_range"]))
    _close(y_ref, y_orc, "long_period/original")
    general["long_period/original/rms"] = np.sqrt(np.mean(np.square(y_ref)))
    general["long_period/original/max"] = np.max(np.abs(y_ref))
    general["long_period/original/dec"] = y_ref[::DECIMATE].copy()
    general["long_period/original/period"] = np.int64(det["period"])
    print("long_period/original period %d pinned" % det["period"])
    for k, v in provenance.items():
        general["provenance/" + k] = np.array(v)
    np.savez_compressed(os.path.join(GOLDEN, "drivers_general.npz"), **general)


def main():
    warnings.simplefilter("ignore")
    ref = reference_shim.load()
    assert ref is not None, "reference not present at /root/reference"
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- the real-audio input ---------------------------------------------------------
    import scipy.io.wavfile

    wav_fs, wav = scipy.io.wavfile.read("/root/reference/audio_file.wav")
    assert wav_fs == FS and wav.dtype == np.int16
    np.savez_compressed(os.path.join(GOLDEN, "audio_file_int16.npz"), pcm=wav, sampling_frequency=wav_fs)
    provenance = {
        "repet_py_sha256": hashlib.sha256(open("/root/reference/repet.py", "rb").read()).hexdigest(),
        "audio_file_sha256": hashlib.sha256(open("/root/reference/audio_file.wav", "rb").read()).hexdigest(),
        "numpy": np.__version__,
    }

    if "--sim-long-only" in sys.argv:
        pin_sim_long(ref)
        return
    if "--general-only" in sys.argv:
        pin_general(ref, provenance)
        return
    # ---- helper-level cases -----------------------------------------------------------
    h = helper_inputs()
    out = {}
    pairs = []
    X = ref._stft(h["signal"], h["window"], h["step"])
    pairs.append(("stft", X, oracle.stft(h["signal"], h["window"], h["step"])))
    pairs.append(("istft", ref._istft(X, h["window"], h["step"]), oracle.istft(X, h["window"], h["step"])))
    V = h["spectrogram"]
    pairs.append(("acorr", ref._acorr(V.T), oracle.acorr(V.T)))
    pairs.append(("beatspectrum", ref._beatspectrum(V), oracle.beatspectrum(V)))
    Bsg = ref._beatspectrogram(V, 60, 30)
    pairs.append(("beatspectrogram", Bsg, oracle.beatspectrogram(V, 60, 30)))
    pairs.append(("periods_1d", np.int64(ref._periods(ref._beatspectrum(V), [3, 50])),
                  np.int64(oracle.periods(oracle.beatspectrum(V), [3, 50]))))
    pairs.append(("periods_2d", ref._periods(Bsg, [3, 50]), oracle.periods(Bsg, [3, 50])))
    pairs.append(("selfsim", ref._selfsimilaritymatrix(V), oracle.selfsimilaritymatrix(V)))
    pairs.append(("sim", ref._similaritymatrix(V, V[:, 10:11]), oracle.similaritymatrix(V, V[:, 10:11])))
    rv, ri = ref._localmaxima(h["vector"], 0.2, 7, 20)
    ov, oi = oracle.localmaxima(h["vector"], 0.2, 7, 20)
    pairs.append(("localmaxima_values", rv, ov))
    pairs.append(("localmaxima_indices", ri, oi))
    S = ref._selfsimilaritymatrix(V)
    rl = ref._indices(S, 0, 9, 12)
    ol = oracle.indices(S, 0, 9, 12)
    pairs.append(("indices_counts", np.array([len(v) for v in rl]), np.array([len(v) for v in ol])))
    pairs.append(("indices_flat", np.concatenate(rl), np.concatenate(ol)))
    for p in (23, 30, 100):  # 300 % 30 == 0: second median block empty (quirk Q9)
        pairs.append(("mask_p%d" % p, ref._mask(V, p), oracle.mask(V, p)))
    pairs.append(("adaptivemask", ref._adaptivemask(V, h["periods_per_frame"], 5),
                  oracle.adaptivemask(V, h["periods_per_frame"], 5)))
    pairs.append(("adaptivemask_order4", ref._adaptivemask(V, h["periods_per_frame"], 4),
                  oracle.adaptivemask(V, h["periods_per_frame"], 4)))
    pairs.append(("simmask", ref._simmask(V, rl), oracle.simmask(V, ol)))
    for name, r, o in pairs:
        _close(r, o, name)
        out[name] = np.asarray(r)
    np.savez_compressed(os.path.join(GOLDEN, "helpers.npz"), **out)
    print("helpers: %d vectors pinned" % len(pairs))

    # ---- driver-level cases -----------------------------------------------------------
    if "--rates-only" not in sys.argv:
        drivers = pin_drivers(ref, DRIVER_CASES, wav)
        for k, v in provenance.items():
            drivers["provenance/" + k] = np.array(v)
        np.savez_compressed(os.path.join(GOLDEN, "drivers.npz"), **drivers)
    rates = pin_drivers(ref, RATE_CASES)
    for k, v in provenance.items():
        rates["provenance/" + k] = np.array(v)
    np.savez_compressed(os.path.join(GOLDEN, "drivers_rates.npz"), **rates)
    pin_sim_long(ref)
    pin_general(ref, provenance)
    sizes = {f: os.path.getsize(os.path.join(GOLDEN, f)) for f in os.listdir(GOLDEN)}
    print("written:", sizes)


if __name__ == "__main__":
    main()
