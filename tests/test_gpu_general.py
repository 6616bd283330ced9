"""GPU: the general-size float64 helper paths (repet_generic.cu) -- _stft / _istft for any window length and step,
_acorr / _beatspectrum for any number of rows (repet.py:1001-1158) -- against the oracle (float64, 1e-10)."""

import numpy as np
import pytest
import scipy.signal.windows

import make_golden
import repet_oracle as oracle
import repet_synth

pytestmark = pytest.mark.gpu

FS = 44100


@pytest.fixture(scope="module")
def repet():
    import repet as module

    module._host.get_handle(0)
    return module


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


@pytest.mark.parametrize("window_length,step_length,samples", [
    (256, 128, 5000),     # the golden helper case of make_golden.helper_inputs (a window no driver uses)
    (256, 64, 5000),      # 75 % overlap: four frames per sample
    (2048, 512, 30000),   # a fast-path length with another hop
    (4096, 2048, 50000),  # 96 kHz window (repet.py:130)
    (8192, 4096, 70000),  # 192 kHz window
    (1000, 300, 12345),   # not a power of two, hop not a divisor: Bluestein
    (255, 100, 4000),     # odd window
    (64, 64, 1000),       # no overlap
    (512, 700, 5000),     # hop longer than the window
])
def test_stft_istft_any_window_and_step(repet, window_length, step_length, samples):
    rng = np.random.default_rng(window_length + step_length)
    x = rng.standard_normal(samples)
    w = scipy.signal.windows.hamming(window_length, sym=False)
    X_ref = oracle.stft(x, w, step_length)
    X = repet._stft(x, w, step_length)
    assert X.shape == X_ref.shape and X.dtype == np.complex128
    assert _rel(X, X_ref) <= 1e-10
    if step_length <= window_length:
        y_ref = oracle.istft(X_ref, w, step_length)
        y = repet._istft(X_ref, w, step_length)
        assert y.shape == y_ref.shape
        assert _rel(y, y_ref) <= 1e-10


def test_stft_golden_helper_vector(repet, golden_helpers):
    """The reference's own outputs for the 256-point helper case (tests/golden/helpers.npz)."""
    h = make_golden.helper_inputs()
    X = repet._stft(h["signal"], h["window"], h["step"])
    assert _rel(X, golden_helpers["stft"]) <= 1e-10
    y = repet._istft(golden_helpers["stft"], h["window"], h["step"])
    assert _rel(y, golden_helpers["istft"]) <= 1e-10


def test_readme_usage_at_96_khz(repet):
    """README.md:79-81 computes abs(repet._stft(mean(x), hamming(N), N/2)) with N from the sampling rate; at 96 kHz
    that is a 4096-point window."""
    fs = 96000
    x = repet_synth.make_clip(31, 3 * fs, 2, fs, 2048).T.astype(np.float64)
    N = pow(2, int(np.ceil(np.log2(0.04 * fs))))
    assert N == 4096
    w = scipy.signal.windows.hamming(N, sym=False)
    spec = np.abs(repet._stft(np.mean(x, axis=1), w, N // 2)[0 : N // 2 + 1, :])
    ref = np.abs(oracle.stft(np.mean(x, axis=1), w, N // 2)[0 : N // 2 + 1, :])
    assert _rel(spec, ref) <= 1e-10


@pytest.mark.parametrize("rows,columns", [(1293, 1025), (3000, 40), (5169, 7), (100, 3)])
def test_acorr_and_beatspectrum_any_size(repet, rows, columns):
    """`repet._beatspectrum(V)` of a 30 s clip (T = 1293 > 1024 frames) used to raise NotImplementedError."""
    rng = np.random.default_rng(rows)
    V = np.abs(rng.standard_normal((columns, rows))) + 0.05 * np.abs(np.sin(np.arange(rows) * 2 * np.pi / 37.0))[None, :]
    b = repet._beatspectrum(V)
    b_ref = oracle.beatspectrum(V)
    assert b.shape == b_ref.shape == (rows,)
    tol = 1e-10 if rows > 1024 else 2e-5  # the small case still goes through the drivers' fp32 beat kernel
    assert _rel(b, b_ref) <= tol
    a = repet._acorr(V.T)
    a_ref = oracle.acorr(V.T)
    assert a.shape == a_ref.shape == (rows, columns)
    assert _rel(a, a_ref) <= (1e-10 if 2 * rows - 1 > 2048 else 2e-5)


def test_beatspectrum_of_a_thirty_second_clip_gives_the_reference_period(repet):
    x = repet_synth.make_clip(11, 30 * FS).astype(np.float64)
    N, w, H = oracle.stft_parameters(FS)
    V = np.stack([np.abs(oracle.stft(x[c], w, H)[: N // 2 + 1]) for c in range(2)], axis=2)
    P = np.power(np.mean(V, axis=2), 2)
    b = repet._beatspectrum(P)
    pr2 = oracle.period_range_frames([1, 10], FS, H)
    assert repet._periods(b, pr2) == oracle.periods(oracle.beatspectrum(P), pr2)
