"""GPU: repet.extended and repet.adaptive (and the helpers they add) against the oracle and the
golden vectors recorded from the reference.  Same bars as tests/test_gpu_parity.py."""

import warnings

import numpy as np
import pytest

import make_golden
import repet_oracle as oracle
import repet_synth

pytestmark = pytest.mark.gpu

FS = 44100
RTOL_SIGNAL = 1e-4


@pytest.fixture(scope="module")
def repet():
    import repet as module

    module._host.get_handle(0)
    return module


def _assert_signal(y, y_ref, what):
    assert y.shape == y_ref.shape, what
    rel = float(np.linalg.norm(np.ravel(y - y_ref)) / max(np.linalg.norm(np.ravel(y_ref)), 1e-300))
    worst = float(np.max(np.abs(y - y_ref)) / max(np.max(np.abs(y_ref)), 1e-300))
    assert rel <= RTOL_SIGNAL and worst <= RTOL_SIGNAL, "%s: rel L2 %.3e, max-abs/max %.3e" % (what, rel, worst)


def _spectrogram(index, seconds):
    x = repet_synth.make_clip(index, int(seconds * FS)).astype(np.float64)
    N, w, H = oracle.stft_parameters(FS)
    return np.stack([np.abs(oracle.stft(x[c], w, H)[: N // 2 + 1]) for c in range(2)], axis=2)


# ---- helpers ---------------------------------------------------------------------------------
def test_beatspectrogram_and_periods_helpers(repet, golden_helpers):
    V = make_golden.helper_inputs()["spectrogram"]  # 33 x 300
    B = repet._beatspectrogram(V, 60, 30)
    ref = golden_helpers["beatspectrogram"]
    assert B.shape == ref.shape
    assert float(np.max(np.abs(B - ref)) / np.max(np.abs(ref))) <= 1e-5
    assert np.all(B[:, 29] == 0) and np.all(B[:, 59] == 0)  # quirk Q3 survives
    # _periods on the REFERENCE's float64 beat spectra: pure argmax, must be bit-exact
    assert np.array_equal(repet._periods(ref, [3, 50]), golden_helpers["periods_2d"])
    assert repet._periods(golden_helpers["beatspectrum"], [3, 50]) == int(golden_helpers["periods_1d"])
    with pytest.raises(ValueError):
        repet._periods(np.zeros(9), [3, 80])


def test_beatspectrogram_full_size(repet):
    V = np.power(np.mean(_spectrogram(12, 14.0), axis=2), 2)
    L, step = 431, 215
    B = repet._beatspectrogram(V, L, step)
    ref = oracle.beatspectrogram(V, L, step)
    assert float(np.max(np.abs(B - ref)) / np.max(np.abs(ref))) <= 1e-5
    pr2 = oracle.period_range_frames([1, 10], FS, 1024)
    assert np.array_equal(repet._periods(B, pr2), oracle.periods(ref, pr2))


@pytest.mark.parametrize("order", [5, 4, 1, 9])
def test_adaptivemask_matches_oracle(repet, order):
    V = _spectrogram(13, 8.0)[:, :, 0]
    T = V.shape[1]
    rng = np.random.default_rng(order)
    periods = rng.integers(3, 120, size=T)
    Vq = V.astype(np.float32).astype(np.float64)
    M_ref = oracle.adaptivemask(Vq, periods, order)
    M = repet._adaptivemask(V, periods, order)
    assert float(np.max(np.abs(M - M_ref))) <= 2e-6


# ---- drivers ---------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["wav_5s", "synth_12s", "synth_21s", "wav_full"])
def test_extended_matches_reference(repet, case, golden_drivers, wav_pcm):
    warnings.simplefilter("ignore")
    x = make_golden.case_input(make_golden.DRIVER_CASES[case], wav_pcm)
    y, periods = repet._host.extended_f64(x, FS, repet._tunables(), return_periods=True)
    key = "%s/extended" % case
    assert np.array_equal(periods, golden_drivers[key + "/periods"]), "segment periods must be bit-exact"
    _assert_signal(y[:: make_golden.DECIMATE], golden_drivers[key + "/dec"], key + " (golden)")
    _assert_signal(y, oracle.extended(x, FS), key + " (oracle)")
    assert np.array_equal(repet.extended(x, FS), y)


def test_extended_other_segment_sizes(repet):
    """Overlap of more than two segments (length 6 s, step 2 s): the sequential cross-fade order matters."""
    x = make_golden.case_input(make_golden.DRIVER_CASES["synth_21s"])
    saved = (repet.segment_length, repet.segment_step, repet.period_range)
    try:
        repet.segment_length, repet.segment_step, repet.period_range = 6, 2, [0.5, 3]
        y, periods = repet._host.extended_f64(x, FS, repet._tunables(), return_periods=True)
        y_ref, det = oracle.extended(x, FS, return_details=True, segment_length=6, segment_step=2, period_range=(0.5, 3))
    finally:
        repet.segment_length, repet.segment_step, repet.period_range = saved
    assert periods.tolist() == det["periods"]
    _assert_signal(y, y_ref, "extended 6 s / 2 s")


@pytest.mark.parametrize("case", ["wav_5s", "synth_12s", "synth_mono_8s", "synth_30s", "wav_full"])
def test_adaptive_matches_reference(repet, case, golden_drivers, wav_pcm):
    warnings.simplefilter("ignore")
    x = make_golden.case_input(make_golden.DRIVER_CASES[case], wav_pcm)
    y, periods = repet._host.adaptive_f64(x, FS, repet._tunables(), return_periods=True)
    key = "%s/adaptive" % case
    assert np.array_equal(periods, golden_drivers[key + "/periods"]), "per-frame periods must be bit-exact"
    _assert_signal(y[:: make_golden.DECIMATE], golden_drivers[key + "/dec"], key + " (golden)")
    _assert_signal(y, oracle.adaptive(x, FS), key + " (oracle)")
    assert np.array_equal(repet.adaptive(x, FS), y)


def test_adaptive_even_filter_order(repet):
    x = make_golden.case_input(make_golden.DRIVER_CASES["synth_12s"])
    saved = repet.filter_order
    try:
        repet.filter_order = 4
        y = repet.adaptive(x, FS)
    finally:
        repet.filter_order = saved
    _assert_signal(y, oracle.adaptive(x, FS, filter_order=4), "adaptive order 4")


def test_batches_match_single_calls(repet):
    audio = repet_synth.make_batch(500, 3, 16 * FS)
    for name in ("extended", "adaptive"):
        background, ints = getattr(repet, name + "_batch")(audio, FS)
        for i in range(audio.shape[0]):
            y_ref, det = getattr(oracle, name)(audio[i].T.astype(np.float64), FS, return_details=True)
            assert np.array_equal(ints[i], np.asarray(det["periods"])), (name, i)
            _assert_signal(background[i].T.astype(np.float64), y_ref, "%s clip %d" % (name, i))
        # shard invariance
        part, part_ints = getattr(repet, name + "_batch")(audio[1:], FS)
        assert np.array_equal(part, background[1:]) and np.array_equal(part_ints, ints[1:])
