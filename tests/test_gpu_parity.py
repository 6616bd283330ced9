"""GPU: the CUDA path, called through the C ABI (ctypes), against the CPU oracle and the golden
vectors recorded from the reference.

Bars (BASELINE.json north_star):
  * integer outputs (periods) bit-exact;
  * float signals: ||y - y_ref||_2 / ||y_ref||_2 <= 1e-4 and max|y - y_ref| <= 1e-4 * max|y_ref|
    against the float64 oracle (the CUDA path computes in fp32; observed ~1e-6).
"""

import warnings

import numpy as np
import pytest

import make_golden
import repet_oracle as oracle
import repet_synth

pytestmark = pytest.mark.gpu

FS = 44100
RTOL_SIGNAL = 1e-4  # stated tolerance of north_star
RTOL_SPECTRUM = 2e-5  # fp32 2048-point transform vs float64


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


# ------------------------------------------------------------------------------------------
# helpers through the ABI
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("number_samples", [2048, 5000, 44100 + 17, 3 * 1024])
def test_stft_matches_oracle(repet, number_samples):
    rng = np.random.default_rng(number_samples)
    x = rng.standard_normal(number_samples)
    N, w, H = oracle.stft_parameters(FS)
    X_ref = oracle.stft(x, w, H)
    X = repet._stft(x, w, H)
    assert X.shape == X_ref.shape and X.dtype == complex
    assert _rel(X, X_ref) <= RTOL_SPECTRUM
    # stereo packing: two different channels through one complex transform
    x2 = np.stack([x, rng.standard_normal(number_samples)])
    half = repet._host.stft_half(x2, w, H)
    for c in range(2):
        ref = oracle.stft(x2[c], w, H)[: N // 2 + 1].T
        assert _rel(half[c], ref) <= RTOL_SPECTRUM, c


def test_stft_power_output(repet):
    x = repet_synth.make_clip(2, 40000).astype(np.float64)
    N, w, H = oracle.stft_parameters(FS)
    half, power = repet._host.stft_half(x, w, H, with_power=True)
    mags = np.stack([np.abs(oracle.stft(x[c], w, H)[: N // 2 + 1]) for c in range(2)], axis=2)
    ref = np.power(np.mean(mags, axis=2), 2).T
    assert _rel(power, ref) <= 5e-5


@pytest.mark.parametrize("number_samples", [4096, 30000])
def test_istft_matches_oracle_and_round_trips(repet, number_samples):
    rng = np.random.default_rng(7 + number_samples)
    x = rng.standard_normal(number_samples)
    N, w, H = oracle.stft_parameters(FS)
    X_ref = oracle.stft(x, w, H)
    y_ref = oracle.istft(X_ref, w, H)
    y = repet._istft(X_ref, w, H)
    assert y.shape == y_ref.shape
    _assert_signal(y, y_ref, "istft")
    # analysis -> synthesis is the identity on the first S samples (COLA)
    back = repet._istft(repet._stft(x, w, H), w, H)[:number_samples]
    _assert_signal(back, x, "round trip")


def test_istft_ignores_antihermitian_part(repet):
    """real(ifft(.)) of the reference only sees the Hermitian part of its input (repet.py:1085)."""
    rng = np.random.default_rng(3)
    N, w, H = oracle.stft_parameters(FS)
    Y = rng.standard_normal((N, 9)) + 1j * rng.standard_normal((N, 9))
    _assert_signal(repet._istft(Y, w, H), oracle.istft(Y, w, H), "istft of arbitrary complex input")


def _spectrogram(index, seconds):
    x = repet_synth.make_clip(index, int(seconds * FS)).astype(np.float64)
    N, w, H = oracle.stft_parameters(FS)
    mags = np.stack([np.abs(oracle.stft(x[c], w, H)[: N // 2 + 1]) for c in range(2)], axis=2)
    return mags


@pytest.mark.parametrize("seconds", [5.0, 20.0])
def test_beatspectrum_matches_oracle(repet, seconds):
    V = np.power(np.mean(_spectrogram(4, seconds), axis=2), 2)
    b_ref = oracle.beatspectrum(V)
    b = repet._beatspectrum(V)
    assert b.shape == b_ref.shape and b.dtype == np.float64
    assert float(np.max(np.abs(b - b_ref)) / np.max(np.abs(b_ref))) <= 1e-5


def test_beatspectrum_small_matrix(repet, golden_helpers):
    """33 x 300 helper vector of the golden set (fewer rows than 1025: zero rows add nothing)."""
    V = make_golden.helper_inputs()["spectrogram"]
    b = repet._beatspectrum(V)
    ref = golden_helpers["beatspectrum"]
    assert float(np.max(np.abs(b - ref)) / np.max(np.abs(ref))) <= 1e-5
    assert repet._host.period_of(V, [3, 50]) == int(golden_helpers["periods_1d"])


@pytest.mark.parametrize("seconds", [4.0, 12.0, 30.0])
def test_period_bit_exact(repet, seconds):
    V = np.power(np.mean(_spectrogram(9, seconds), axis=2), 2)
    pr2 = oracle.period_range_frames([1, 10], FS, 1024)
    assert repet._host.period_of(V, pr2) == int(oracle.periods(oracle.beatspectrum(V), pr2))


@pytest.mark.parametrize("period", [44, 45, 100, 129, 200, 431])
def test_mask_matches_oracle(repet, period):
    V = _spectrogram(6, 15.0)[:, :, 0]  # T = 648: n = 2..15 values per median
    M_ref = oracle.mask(V.astype(np.float32).astype(np.float64), period)
    M = repet._mask(V, period)
    assert M.shape == M_ref.shape
    assert float(np.max(np.abs(M - M_ref))) <= 2e-6  # masks live in [0, 1]


def test_mask_exact_multiple_and_long_median(repet):
    V = _spectrogram(8, 9.0)[:, :, 1]
    T = V.shape[1]
    Vq = V.astype(np.float32).astype(np.float64)
    for period in (T // 4 if T % 4 == 0 else 97, 5, 3):  # 5 and 3 -> n > 32: rank-selection path
        assert float(np.max(np.abs(repet._mask(V, period) - oracle.mask(Vq, period)))) <= 2e-6, period
    V2 = V[:, : (T // 50) * 50]  # T an exact multiple of the period (quirk Q9, second block empty)
    assert float(np.max(np.abs(repet._mask(V2, 50) - oracle.mask(V2.astype(np.float32).astype(np.float64), 50)))) <= 2e-6


# ------------------------------------------------------------------------------------------
# repet.original through the reference's own calling convention
# ------------------------------------------------------------------------------------------
ORIGINAL_CASES = ["wav_5s", "synth_12s", "synth_21s", "synth_mono_8s", "synth_30s", "wav_full"]


@pytest.mark.parametrize("case", ORIGINAL_CASES)
def test_original_matches_reference(repet, case, golden_drivers, wav_pcm):
    warnings.simplefilter("ignore")
    spec = make_golden.DRIVER_CASES[case]
    x = make_golden.case_input(spec, wav_pcm)
    y, period = repet._host.original_f64(x, FS, repet._tunables(), return_period=True)
    key = "%s/original" % case
    assert y.dtype == np.float64 and y.shape == x.shape
    assert period == int(golden_drivers[key + "/period"]), "period must be bit-exact"
    # against the reference's recorded samples ...
    dec = golden_drivers[key + "/dec"]
    _assert_signal(y[:: make_golden.DECIMATE], dec, key + " (golden)")
    # ... and against the oracle on every sample
    _assert_signal(y, oracle.original(x, FS), key + " (oracle)")
    # the public entry point is the same call
    assert np.array_equal(repet.original(x, FS), y)


def test_original_reads_module_tunables_at_call_time(repet):
    x = make_golden.case_input(make_golden.DRIVER_CASES["synth_12s"])
    saved = (repet.period_range, repet.cutoff_frequency)
    try:
        repet.period_range = [1, 2]
        repet.cutoff_frequency = 300
        y = repet.original(x, FS)
        y_ref = oracle.original(x, FS, period_range=(1, 2), cutoff_frequency=300)
    finally:
        repet.period_range, repet.cutoff_frequency = saved
    _assert_signal(y, y_ref, "tunables")


def test_original_error_behaviour(repet):
    with pytest.raises(ValueError):
        repet.original(np.zeros(44100), FS)  # 1-D input: shape unpack fails (quirk Q17)
    with pytest.raises(ValueError):
        repet.original(np.full((2 * FS, 2), 0.01), FS)  # too short: argmax of an empty sequence
    # the input is never mutated
    x = make_golden.case_input(make_golden.DRIVER_CASES["synth_mono_8s"])
    before = x.copy()
    repet.original(x, FS)
    assert np.array_equal(x, before)


# ------------------------------------------------------------------------------------------
# batch API: host buffers, device buffers, shard invariance
# ------------------------------------------------------------------------------------------
def _batch(number_clips, seconds, first=40):
    return repet_synth.make_batch(first, number_clips, int(seconds * FS))


def test_original_batch_matches_oracle_per_clip(repet):
    audio = _batch(5, 8.0)
    background, periods = repet.original_batch(audio, FS)
    assert background.shape == audio.shape and background.dtype == np.float32
    for i in range(audio.shape[0]):
        y_ref, det = oracle.original(audio[i].T.astype(np.float64), FS, return_details=True)
        assert int(periods[i]) == det["period"], i
        _assert_signal(background[i].T.astype(np.float64), y_ref, "clip %d" % i)


def test_original_batch_is_shard_invariant(repet):
    """Clips are independent: any split of the batch gives byte-identical results."""
    audio = _batch(6, 6.0, first=70)
    whole, periods = repet.original_batch(audio, FS)
    handle = repet._host.get_handle(0)
    handle.set_workspace_limit(64 << 20)  # forces one-clip chunks
    try:
        chunked, periods_chunked = repet.original_batch(audio, FS)
    finally:
        handle.set_workspace_limit(0)
    assert np.array_equal(whole, chunked) and np.array_equal(periods, periods_chunked)
    for lo, hi in ((0, 2), (2, 6)):
        part, part_periods = repet.original_batch(audio[lo:hi], FS)
        assert np.array_equal(part, whole[lo:hi]) and np.array_equal(part_periods, periods[lo:hi])


def test_original_batch_device_pointers(repet):
    torch = pytest.importorskip("torch")
    audio = _batch(3, 7.0, first=90)
    host_result, host_periods = repet.original_batch(audio, FS)
    device = torch.device("cuda:0")
    audio_dev = torch.from_numpy(audio).to(device)
    out_dev = torch.empty_like(audio_dev)
    periods_dev = torch.zeros(audio.shape[0], dtype=torch.int32, device=device)
    handle = repet._host.get_handle(0)
    stream = torch.cuda.current_stream(device)
    handle.set_stream(stream.cuda_stream)
    try:
        repet._host.original_batch_device(
            audio_dev.data_ptr(), out_dev.data_ptr(), audio.shape[0], audio.shape[1], audio.shape[2], FS,
            repet._tunables(), handle=handle, periods_ptr=periods_dev.data_ptr(),
        )
        stream.synchronize()
    finally:
        handle.set_stream(None)
    assert np.array_equal(out_dev.cpu().numpy(), host_result)
    assert np.array_equal(periods_dev.cpu().numpy(), host_periods)


def test_full_size_clip_properties(repet):
    """BASELINE config 2 clip size (30 s stereo), size-independent properties: a clip's result
    does not depend on its batch neighbours; scaling the input by 2 scales the output by 2 and
    keeps the period (every stage is homogeneous; eps only matters for bins below 1e-9);
    swapping the channels swaps the outputs (the beat spectrum uses the channel mean)."""
    audio = _batch(3, 30.0, first=200)
    background, periods = repet.original_batch(audio, FS)
    alone, period_alone = repet.original_batch(audio[1:2], FS)
    assert np.array_equal(alone[0], background[1]) and period_alone[0] == periods[1]
    doubled, periods_doubled = repet.original_batch(2.0 * audio, FS)
    assert np.array_equal(periods_doubled, periods)
    assert _rel(doubled, 2.0 * background) <= 1e-6
    swapped, periods_swapped = repet.original_batch(np.ascontiguousarray(audio[:, ::-1, :]), FS)
    assert np.array_equal(periods_swapped, periods)
    assert _rel(swapped[:, ::-1, :], background) <= 1e-5


def test_original_long_clip_blocked_beat_transform(repet):
    """Clips longer than one 2048-point beat transform (T + max lag > 2048 frames, i.e. > ~37 s)
    take the blocked cross-spectrum path; medians over more than 32 segments take the
    rank-selection path.  90 s of audio: T = 3877."""
    x = repet_synth.make_clip(77, 90 * FS).T.astype(np.float64)
    y, period = repet._host.original_f64(x, FS, repet._tunables(), return_period=True)
    y_ref, det = oracle.original(x, FS, return_details=True)
    assert period == det["period"]
    _assert_signal(y, y_ref, "original, 90 s")
    # a short period range makes ceil(T / p) > 32
    saved = repet.period_range
    try:
        repet.period_range = [0.5, 2]
        y2, period2 = repet._host.original_f64(x, FS, repet._tunables(), return_period=True)
    finally:
        repet.period_range = saved
    y2_ref, det2 = oracle.original(x, FS, return_details=True, period_range=(0.5, 2))
    assert period2 == det2["period"] and -(-3877 // period2) > 32
    _assert_signal(y2, y2_ref, "original, 90 s, short periods")


def test_original_batch_pcm16_matches_float_path(repet, wav_pcm):
    """int16 PCM in WAV order, normalised on the device exactly as repet.wavread does (x / 2^15)."""
    pcm = np.stack([wav_pcm[0 : 6 * FS], wav_pcm[8 * FS : 14 * FS]])  # (2, S, 2) int16
    background, periods = repet.original_batch_pcm16(pcm, FS)
    audio = np.ascontiguousarray(np.transpose(pcm / 32768.0, (0, 2, 1)), dtype=np.float32)
    ref_background, ref_periods = repet.original_batch(audio, FS)
    assert np.array_equal(periods, ref_periods)
    assert np.array_equal(background, ref_background)  # int16 / 2^15 is exact in fp32


def test_original_periods_bit_exact_over_many_clips(repet):
    """24 of the benchmark's 30 s clips: every period equals the float64 oracle's.  The beat spectrum's
    relative top-1/top-2 gap is printed: the float64 period search has to resolve the smallest one."""
    audio = _batch(24, 30.0, first=0)
    _, periods = repet.original_batch(audio, FS)
    pr2 = oracle.period_range_frames([1, 10], FS, 1024)
    gaps = []
    for i in range(audio.shape[0]):
        x = audio[i].T.astype(np.float64)
        _, det = oracle.original(x, FS, return_details=True)
        assert int(periods[i]) == det["period"], "clip %d: gpu %d, oracle %d" % (i, periods[i], det["period"])
        b = det["beat_spectrum"]
        hi = min(int(pr2[1]), len(b) // 3)
        window = np.sort(b[int(pr2[0]) : hi])
        gaps.append(float((window[-1] - window[-2]) / window[-1]))
    print("relative top-2 gaps of the beat spectrum: min %.3e, median %.3e" % (min(gaps), float(np.median(gaps))))


def test_period_certification_path(repet):
    """Force the float64 re-evaluation of near-tied lags (window widened from 100 ppm to 20 %): the
    certified periods must equal the oracle's, i.e. the exact path agrees with the fast one."""
    audio = _batch(6, 20.0, first=300)
    _, fast = repet.original_batch(audio, FS)
    repet._host.set_tuning(cert_rel_ppm=200000)
    try:
        background, certified = repet.original_batch(audio, FS)
    finally:
        repet._host.set_tuning(cert_rel_ppm=0)
    assert np.array_equal(fast, certified)
    for i in range(audio.shape[0]):
        _, det = oracle.original(audio[i].T.astype(np.float64), FS, return_details=True)
        assert int(certified[i]) == det["period"], i


def test_adaptive_period_certification_path(repet):
    """The same for the per-segment periods of the adaptive REPET (segments are windows of the clip's power
    spectrogram, zero outside the clip): with the certification window widened to 20 % most segments take the
    float64 re-evaluation, and every frame's period must still equal the oracle's."""
    audio = _batch(3, 24.0, first=340)
    _, fast = repet.adaptive_batch(audio, FS)
    repet._host.set_tuning(cert_rel_ppm=200000)
    try:
        _, certified = repet.adaptive_batch(audio, FS)
    finally:
        repet._host.set_tuning(cert_rel_ppm=0)
    assert np.array_equal(fast, certified)
    for i in range(audio.shape[0]):
        _, det = oracle.adaptive(audio[i].T.astype(np.float64), FS, return_details=True)
        assert np.array_equal(certified[i], det["periods"]), i
