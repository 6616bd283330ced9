"""GPU: the by-products every documented use of the reference computes right after the call
(README.md:64-81 of the reference): foreground = audio - background and the display spectrograms
abs(_stft(mean(x, axis=1)))[0:F] of mixture, background and foreground -- produced on the device by
`repet.separate`, checked against the same expressions on the float64 oracle."""

import ctypes

import numpy as np
import pytest

import make_golden
import repet_oracle as oracle

pytestmark = pytest.mark.gpu

FS = 44100
RTOL = 1e-4


@pytest.fixture(scope="module")
def repet():
    import repet as module

    module._host.get_handle(0)
    return module


def _display_spectrogram(signal, fs):
    N, w, H = oracle.stft_parameters(fs)
    return np.abs(oracle.stft(np.mean(signal, axis=1), w, H)[0 : N // 2 + 1, :])  # README.md:79


def _rel(a, b):
    return float(np.linalg.norm(np.ravel(a - b)) / max(np.linalg.norm(np.ravel(b)), 1e-300))


@pytest.mark.parametrize("method,case", [("original", "synth_12s"), ("adaptive", "synth_mono_8s"), ("sim", "synth_12s"),
                                         ("extended", "synth_21s"), ("simonline", "synth_12s")])
def test_separate_matches_the_documented_expressions(repet, method, case):
    x = make_golden.case_input(make_golden.DRIVER_CASES[case])
    out = repet.separate(x, FS, method=method)
    y_ref = getattr(oracle, method)(x, FS)
    assert out["background"].shape == x.shape and out["foreground"].shape == x.shape
    assert np.array_equal(out["background"], getattr(repet, method)(x, FS))  # same path as the plain call
    assert _rel(out["background"], y_ref) <= RTOL
    # the foreground is formed in fp32 on the device: error relative to the mixture, not to itself
    assert float(np.max(np.abs(out["foreground"] - (x - y_ref)))) <= RTOL * float(np.max(np.abs(x)))
    for name, signal in (("audio_spectrogram", x), ("background_spectrogram", y_ref), ("foreground_spectrogram", x - y_ref)):
        ref = _display_spectrogram(signal, FS)
        got = out[name]
        assert got.shape == ref.shape, name
        assert _rel(got, ref) <= RTOL, "%s: rel L2 %.3e" % (name, _rel(got, ref))
        assert float(np.max(np.abs(got - ref))) <= RTOL * float(np.max(ref)), name


def test_separate_integer_outputs(repet):
    x = make_golden.case_input(make_golden.DRIVER_CASES["synth_12s"])
    out = repet.separate(x, FS, method="original", spectrograms=False)
    _, det = oracle.original(x, FS, return_details=True)
    assert int(out["integers"][0]) == det["period"]
    assert "audio_spectrogram" not in out
    with pytest.raises(ValueError):
        repet.separate(x, FS, method="median")


def test_spectrogram_and_db_helpers(repet):
    x = make_golden.case_input(make_golden.DRIVER_CASES["synth_12s"])
    ref = _display_spectrogram(x, FS)
    got = repet.spectrogram(x, FS)
    assert got.shape == ref.shape and _rel(got, ref) <= RTOL
    assert np.allclose(repet.spectrogram_db(ref), 20 * np.log10(ref))  # repet.py:982


def test_device_resident_spectrogram_and_foreground(repet):
    """The batch entry points on device pointers: spectrograms of 3 stereo clips and a foreground pass."""
    torch = pytest.importorskip("torch")
    import repet_synth

    audio = repet_synth.make_batch(900, 3, 5 * FS + 17)
    dev = torch.from_numpy(audio).cuda()
    handle = repet._host.get_handle(0)
    params, _ = repet._host.derive_params(FS, repet._tunables())
    frames = handle.lib.repet_spectrogram_frames(ctypes.byref(params), audio.shape[2])
    pitch = handle.lib.repet_spectrogram_pitch(ctypes.byref(params))
    spec = torch.empty((3, frames, pitch), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    assert repet._host.spectrogram_batch_device(dev.data_ptr(), spec.data_ptr(), 3, 2, audio.shape[2], FS,
                                                repet._tunables(), handle=handle) == (frames, pitch)
    handle.synchronize()
    got = spec.cpu().numpy()
    for i in range(3):
        ref = _display_spectrogram(audio[i].T.astype(np.float64), FS)
        assert _rel(got[i, :, :1025].T, ref) <= RTOL
    background = 0.25 * dev
    foreground = torch.empty_like(dev)
    torch.cuda.synchronize()
    repet._host.foreground_device(dev.data_ptr(), background.data_ptr(), foreground.data_ptr(), dev.numel(), handle=handle)
    handle.synchronize()
    assert torch.equal(foreground, dev - background)
