"""CPU: the oracle (oracle/repet_oracle.py) against the golden vectors that
oracle/make_golden.py recorded from the UNMODIFIED reference repet.py.

Integers bit-exact; floats to 1e-12 relative (they are bit-identical on the NumPy that
generated them; the slack only covers a different BLAS/pocketfft build)."""

import warnings

import numpy as np
import pytest

import make_golden
import repet_oracle as oracle

FS = 44100
RTOL = 1e-12


def _close(actual, expected):
    actual = np.asarray(actual)
    expected = np.asarray(expected)
    assert actual.shape == expected.shape
    if expected.dtype.kind in "iu":
        assert np.array_equal(actual, expected)
        return
    scale = max(float(np.max(np.abs(expected))) if expected.size else 0.0, 1e-300)
    assert float(np.max(np.abs(actual - expected))) / scale <= RTOL


def test_helper_vectors(golden_helpers):
    g = golden_helpers
    h = make_golden.helper_inputs()
    X = oracle.stft(h["signal"], h["window"], h["step"])
    _close(X, g["stft"])
    _close(oracle.istft(X, h["window"], h["step"]), g["istft"])
    V = h["spectrogram"]
    _close(oracle.acorr(V.T), g["acorr"])
    _close(oracle.beatspectrum(V), g["beatspectrum"])
    B = oracle.beatspectrogram(V, 60, 30)
    _close(B, g["beatspectrogram"])
    # quirk Q3: column i+step-1 of the beat spectrogram stays zero
    assert np.all(B[:, 29] == 0) and np.all(B[:, 59] == 0) and np.any(B[:, 28] != 0)
    _close(np.int64(oracle.periods(oracle.beatspectrum(V), [3, 50])), g["periods_1d"])
    _close(oracle.periods(B, [3, 50]), g["periods_2d"])
    _close(oracle.selfsimilaritymatrix(V), g["selfsim"])
    _close(oracle.similaritymatrix(V, V[:, 10:11]), g["sim"])
    values, index = oracle.localmaxima(h["vector"], 0.2, 7, 20)
    _close(values, g["localmaxima_values"])
    _close(index, g["localmaxima_indices"])
    assert not set(range(50, 55)) & set(index.tolist())  # plateau has no strict maximum (Q7)
    lists = oracle.indices(oracle.selfsimilaritymatrix(V), 0, 9, 12)
    _close(np.array([len(v) for v in lists]), g["indices_counts"])
    _close(np.concatenate(lists), g["indices_flat"])
    for p in (23, 30, 100):
        _close(oracle.mask(V, p), g["mask_p%d" % p])
    _close(oracle.adaptivemask(V, h["periods_per_frame"], 5), g["adaptivemask"])
    _close(oracle.adaptivemask(V, h["periods_per_frame"], 4), g["adaptivemask_order4"])
    _close(oracle.simmask(V, lists), g["simmask"])


def test_periods_quirks():
    # Q1: argmax + 1 + lo ; Q2: upper bound min(hi, n//3), exclusive
    b = np.zeros(100)
    b[10] = 1.0
    assert oracle.periods(b, [3, 30]) == 11
    b = np.zeros(90)
    b[29] = 1.0
    b[30] = 5.0  # lag 30 == 90//3 is excluded
    assert oracle.periods(b, [3, 80]) == 30
    assert np.array_equal(oracle.periods(np.zeros((90, 4)), [3, 80]), np.full(4, 4))  # all-zero column -> lo+1 (Q3)
    with pytest.raises(ValueError):
        oracle.periods(np.zeros(9), [3, 80])  # empty range (Q17)


# (case, function) pairs cheap enough for the CPU suite; the rest run in the gpu suite's
# oracle comparisons and were all asserted when the goldens were made
FAST = [
    ("wav_5s", "original"), ("wav_5s", "extended"), ("wav_5s", "adaptive"), ("wav_5s", "sim"),
    ("synth_12s", "original"), ("synth_12s", "extended"), ("synth_12s", "adaptive"), ("synth_12s", "sim"),
    ("synth_12s", "simonline"), ("synth_21s", "extended"), ("synth_mono_8s", "original"),
    ("synth_mono_8s", "adaptive"), ("synth_mono_8s", "sim"), ("wav_full", "original"),
    ("wav_full", "extended"), ("wav_full", "adaptive"), ("synth_30s", "original"),
]


@pytest.mark.parametrize("case,fn", FAST)
def test_driver_vectors(case, fn, golden_drivers, wav_pcm):
    warnings.simplefilter("ignore")
    g = golden_drivers
    spec = make_golden.DRIVER_CASES[case]
    x = make_golden.case_input(spec, wav_pcm)
    y, det = getattr(oracle, fn)(x, FS, return_details=True)
    key = "%s/%s" % (case, fn)
    assert y.shape == x.shape and y.dtype == np.float64
    _close(y[:: make_golden.DECIMATE], g[key + "/dec"])
    _close(np.sqrt(np.mean(np.square(y))), g[key + "/rms"])
    if fn == "original":
        assert det["period"] == int(g[key + "/period"])
    elif fn in ("extended", "adaptive"):
        _close(np.asarray(det["periods"]), g[key + "/periods"])
    elif fn in ("sim", "simonline"):
        _close(np.array([len(v) for v in det["indices"]]), g[key + "/index_counts"])
        _close(np.concatenate(det["indices"]), g[key + "/index_flat"])


RATE_FAST = [(case, fn) for case, spec in make_golden.RATE_CASES.items() for fn in spec["functions"]]


@pytest.mark.parametrize("case,fn", RATE_FAST)
def test_driver_vectors_other_sampling_rates(case, fn, golden_rates):
    """Window lengths 512 / 1024 and the 48 kHz hop mapping (repet.py:130, quirk Q16)."""
    warnings.simplefilter("ignore")
    g = golden_rates
    spec = make_golden.RATE_CASES[case]
    x = make_golden.case_input(spec)
    y, det = getattr(oracle, fn)(x, make_golden.case_fs(spec), return_details=True)
    key = "%s/%s" % (case, fn)
    _close(y[:: make_golden.DECIMATE], g[key + "/dec"])
    if fn == "original":
        assert det["period"] == int(g[key + "/period"])
    elif fn in ("extended", "adaptive"):
        _close(np.asarray(det["periods"]), g[key + "/periods"])
    elif fn in ("sim", "simonline"):
        _close(np.array([len(v) for v in det["indices"]]), g[key + "/index_counts"])
        _close(np.concatenate(det["indices"]), g[key + "/index_flat"])


def test_known_answers_cfg1(golden_drivers):
    """The known answers SURVEY.md section 8(c) recorded for BASELINE config 1."""
    g = golden_drivers
    assert int(g["wav_full/original/period"]) == 286
    assert float(g["wav_full/original/rms"]) == pytest.approx(0.11790343090999283, rel=1e-12)
    assert g["wav_full/extended/periods"].tolist() == [46, 45, 45]
    assert float(g["wav_full/adaptive/rms"]) == pytest.approx(0.09515059059406461, rel=1e-12)
    periods = g["wav_full/adaptive/periods"]
    assert periods[[214, 429, 644, 859]].tolist() == [44, 44, 44, 44]  # quirk Q3
    counts = g["wav_full/sim/index_counts"]
    assert (counts.min(), counts.max()) == (5, 15) and len(counts) == 992
    assert g["wav_full/sim/index_flat"][:7].tolist() == [0, 568, 309, 853, 744, 674, 166]
    assert float(g["wav_full/simonline/rms"]) == pytest.approx(0.09540832134384586, rel=1e-12)


def test_simonline_too_short_raises():
    x = np.zeros((44100 * 5, 2)) + 0.01
    with pytest.raises(ValueError):
        oracle.simonline(x, FS)


def test_sim_long_hard_list():
    """The one list of the 2-minute REPET-SIM track that an fp32 front end gets wrong (two similarities tied
    to ~1e-8): the float64 oracle reproduces the reference's list for that frame (tests/golden/sim_long.npz)."""
    import os

    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "sim_long.npz"))
    spec = make_golden.SIM_LONG
    x = make_golden.sim_long_input()
    N, w, H = oracle.stft_parameters(44100)
    V = np.mean(np.stack([np.abs(oracle.stft(x[:, c], w, H)[: N // 2 + 1]) for c in range(2)], axis=2), axis=2)
    assert V.shape[1] == len(golden["counts"])
    A = V / np.sqrt(np.sum(np.power(V, 2), axis=0))
    column = np.matmul(A.T, A[:, spec["hard_frame"]])
    _, idx = oracle.localmaxima(column, 0, int(round(44100 / H)), 100)
    assert np.array_equal(idx, golden["hard_list"])


def test_long_goldens_against_the_oracle():
    """The at-size goldens (oracle/make_golden_long.py, recorded from the unmodified reference): the 1-minute
    adaptive case is cheap enough to re-derive with the oracle on every CPU run; the others are checked for shape."""
    import os

    import make_golden_long

    golden_dir = os.path.join(os.path.dirname(__file__), "golden")
    g = np.load(os.path.join(golden_dir, "long_adaptive_1min.npz"))
    x = make_golden_long.long_input("adaptive_1min")
    y, det = oracle.adaptive(x, 44100, return_details=True)
    assert np.array_equal(np.asarray(det["periods"]), g["periods"].astype(np.int64))
    _close(y[:: int(g["decimate"])], g["dec"])
    for case, spec in make_golden_long.CASES.items():
        path = os.path.join(golden_dir, "long_%s.npz" % case)
        if not os.path.isfile(path):
            continue
        data = np.load(path)
        assert int(data["samples"]) == spec["samples"]
        frames = -(-spec["samples"] // 1024) + 1
        if spec["fn"] == "adaptive":
            assert data["periods"].shape == (frames,)
        elif spec["fn"] == "sim":
            assert data["counts"].shape == (frames,) and int(data["total"]) == int(data["counts"].astype(np.int64).sum())
        elif spec["fn"] == "simonline":
            online_frames = (spec["samples"] - 2048 + 1023) // 1024 + 1
            assert data["counts"].shape == (online_frames - int(data["first_frame"]),)
