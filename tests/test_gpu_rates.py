"""GPU: the drivers and helpers at other sampling rates.

The reference picks its STFT window from the sampling rate, N = 2^ceil(log2(0.04 fs)) (repet.py:130):
512 points at 8 kHz, 1024 at 16 / 22.05 kHz, 2048 at 44.1 / 48 kHz.  The library carries one
instantiation of every kernel per window length; these tests run each of them against the goldens
recorded from the reference (tests/golden/drivers_rates.npz) and against the oracle.
Bars as in test_gpu_parity.py: integers bit-exact, signals within 1e-4.
"""

import warnings

import numpy as np
import pytest

import make_golden
import repet_oracle as oracle
import repet_synth

pytestmark = pytest.mark.gpu

RTOL_SIGNAL = 1e-4
RTOL_SPECTRUM = 2e-5
RATES = [8000, 11025, 16000, 22050, 32000, 48000]


@pytest.fixture(scope="module")
def repet():
    import repet as module

    module._host.get_handle(0)  # raises if the CUDA library or device is missing: no fallback
    return module


def _rel(a, b):
    return float(np.linalg.norm(np.ravel(a - b)) / max(np.linalg.norm(np.ravel(b)), 1e-300))


def _assert_signal(y, y_ref, what):
    assert y.shape == y_ref.shape, what
    rel = _rel(y, y_ref)
    worst = float(np.max(np.abs(y - y_ref)) / max(np.max(np.abs(y_ref)), 1e-300))
    assert rel <= RTOL_SIGNAL and worst <= RTOL_SIGNAL, "%s: rel L2 %.3e, max-abs/max %.3e" % (what, rel, worst)


def _assert_lists(lists, counts_ref, flat_ref, what, first=0):
    counts = np.array([len(v) for v in lists[first:]])
    assert np.array_equal(counts, counts_ref), what
    flat = np.concatenate(lists[first:]) if len(lists) > first else np.zeros(0, dtype=np.int64)
    assert np.array_equal(flat, flat_ref), what


@pytest.mark.parametrize("fs", RATES)
def test_stft_istft_every_window_length(repet, fs):
    rng = np.random.default_rng(fs)
    x = rng.standard_normal(3 * fs // 2 + 77)
    N, w, H = oracle.stft_parameters(fs)
    X_ref = oracle.stft(x, w, H)
    X = repet._stft(x, w, H)
    assert X.shape == X_ref.shape
    assert _rel(X, X_ref) <= RTOL_SPECTRUM
    y_ref = oracle.istft(X_ref, w, H)
    y = repet._istft(X_ref, w, H)
    assert y.shape == y_ref.shape
    assert _rel(y, y_ref) <= RTOL_SPECTRUM
    # stereo packing
    x2 = np.stack([x, rng.standard_normal(len(x))])
    half, power = repet._host.stft_half(x2, w, H, with_power=True)
    mags = []
    for c in range(2):
        ref = oracle.stft(x2[c], w, H)[: N // 2 + 1].T
        assert _rel(half[c], ref) <= RTOL_SPECTRUM, c
        mags.append(np.abs(ref))
    assert _rel(power, np.power(np.mean(np.stack(mags, axis=2), axis=2), 2)) <= 5e-5


@pytest.mark.parametrize("case,fn", [(c, f) for c, s in make_golden.RATE_CASES.items() for f in s["functions"]])
def test_drivers_match_reference(repet, case, fn, golden_rates):
    warnings.simplefilter("ignore")
    spec = make_golden.RATE_CASES[case]
    fs = make_golden.case_fs(spec)
    x = make_golden.case_input(spec)
    key = "%s/%s" % (case, fn)
    g = golden_rates
    tun = repet._tunables()
    if fn == "original":
        y, period = repet._host.original_f64(x, fs, tun, return_period=True)
        assert int(period) == int(g[key + "/period"]), key
    elif fn == "extended":
        y, periods = repet._host.extended_f64(x, fs, tun, return_periods=True)
        assert np.array_equal(periods, g[key + "/periods"]), key
    elif fn == "adaptive":
        y, periods = repet._host.adaptive_f64(x, fs, tun, return_periods=True)
        assert np.array_equal(periods, g[key + "/periods"]), key
    elif fn == "sim":
        y, lists = repet._host.sim_f64(x, fs, tun, return_indices=True)
        _assert_lists(lists, g[key + "/index_counts"], g[key + "/index_flat"], key)
    else:
        y, lists = repet._host.simonline_f64(x, fs, tun, return_indices=True)
        first = int(g[key + "/first_frame"])
        _assert_lists(lists, g[key + "/index_counts"], g[key + "/index_flat"], key, first=first)
    _assert_signal(y[:: make_golden.DECIMATE], g[key + "/dec"], key + " (golden)")
    assert np.array_equal(getattr(repet, fn)(x, fs), y)


@pytest.mark.parametrize("fs", [8000, 16000])
def test_mask_helpers_other_bin_counts(repet, fs):
    """_mask / _adaptivemask / _simmask on 257- and 513-bin spectrograms."""
    N, w, H = oracle.stft_parameters(fs)
    x = repet_synth.make_clip(31, 6 * fs, 1, fs, H)[0].astype(np.float64)
    V = np.abs(oracle.stft(x, w, H)[: N // 2 + 1])
    T = V.shape[1]
    for period in (7, 20):
        assert _rel(repet._mask(V, period), oracle.mask(V, period)) <= 2e-5
    rng = np.random.default_rng(fs)
    per = rng.integers(3, 25, size=T)
    assert _rel(repet._adaptivemask(V, per, 5), oracle.adaptivemask(V, per, 5)) <= 2e-5
    lists = [np.sort(rng.choice(T, size=rng.integers(1, 40), replace=False)) for _ in range(T)]
    assert _rel(repet._simmask(V, lists), oracle.simmask(V, lists)) <= 2e-5


@pytest.mark.parametrize("fs", [8000, 16000])
def test_batch_other_sampling_rates(repet, fs):
    """original_batch over several clips equals per-clip calls; a rate switch on one handle works."""
    N, w, H = oracle.stft_parameters(fs)
    S = 9 * fs + 11
    audio = np.stack([repet_synth.make_clip(700 + i, S, 2, fs, H) for i in range(5)])
    background, periods = repet.original_batch(audio, fs)
    for i in range(audio.shape[0]):
        y_ref, det = oracle.original(audio[i].T.astype(np.float64), fs, return_details=True)
        assert int(periods[i]) == det["period"], i
        _assert_signal(background[i].T.astype(np.float64), y_ref, "clip %d" % i)
    # back to 44.1 kHz on the same handle
    x = make_golden.case_input(make_golden.DRIVER_CASES["synth_mono_8s"])
    y_ref, det = oracle.original(x, 44100, return_details=True)
    _assert_signal(repet.original(x, 44100), y_ref, "44.1 kHz after %d Hz" % fs)


def test_window_lengths_outside_the_fast_kernels_take_the_general_path(repet):
    """Window lengths other than 512 / 1024 / 2048 used to raise NotImplementedError; they now run on the general
    float64 device path (tests/test_gpu_general.py holds the golden cases at 96 and 192 kHz)."""
    fs = 4000  # N = 256
    x = repet_synth.make_clip(810, 9 * fs, 2, fs, 128).T.astype(np.float64)
    y_ref, det = oracle.original(x, fs, return_details=True)
    y, period = repet._host.original_f64(x, fs, repet._tunables(), return_period=True)
    assert period == det["period"]
    _assert_signal(y, y_ref, "4 kHz (256-point window)")
