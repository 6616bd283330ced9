"""GPU: edge cases of the separation path -- digital silence inside a clip (soft mask of 0/0, quirk Q10;
0/0 = NaN column normalisation of REPET-SIM, quirk Q18, which the reference propagates into its output),
an all-silent clip, an empty batch, a clip a few samples long, odd lengths.  The oracle
(bit-identical to the reference on every golden case) says what must come out, NaN positions included."""

import warnings

import numpy as np
import pytest

import repet_oracle as oracle
import repet_synth

pytestmark = pytest.mark.gpu

FS = 44100
RTOL = 1e-4


@pytest.fixture(scope="module")
def repet():
    import repet as module

    module._host.get_handle(0)
    return module


def _clip_with_silence():
    x = repet_synth.make_clip(77, 12 * FS).T.astype(np.float64)
    x[5 * FS : 7 * FS] = 0.0  # two seconds of digital silence: ~85 all-zero frames
    return x


def _assert_same(y, y_ref, what):
    assert y.shape == y_ref.shape, what
    nan_ref = np.isnan(y_ref)
    assert np.array_equal(np.isnan(y), nan_ref), "%s: NaN pattern differs (%d vs %d NaNs)" % (
        what, int(np.isnan(y).sum()), int(nan_ref.sum()))
    a, b = y[~nan_ref], y_ref[~nan_ref]
    if b.size:
        scale = max(float(np.max(np.abs(b))), 1e-300)
        assert float(np.max(np.abs(a - b))) <= RTOL * scale, what
        assert float(np.linalg.norm(a - b)) <= RTOL * max(float(np.linalg.norm(b)), 1e-300), what


@pytest.mark.parametrize("fn", ["original", "extended", "adaptive"])
def test_silent_stretch_period_methods(repet, fn):
    warnings.simplefilter("ignore")
    x = _clip_with_silence()
    y_ref, det = getattr(oracle, fn)(x, FS, return_details=True)
    y = getattr(repet, fn)(x, FS)
    _assert_same(y, y_ref, fn)
    assert not np.isnan(y).any()  # V = 0 gives mask 1, not NaN (quirk Q10)
    if fn == "original":
        _, period = repet._host.original_f64(x, FS, repet._tunables(), return_period=True)
        assert period == det["period"]


def test_silent_stretch_sim_propagates_nan_like_the_reference(repet):
    """All-zero frames normalise to 0/0 = NaN (quirk Q18); a NaN similarity is never a maximum and blocks its
    neighbours (quirk Q7); frames left with an empty list get a NaN model -- the reference's output holds NaN
    there, and so must ours, at the same samples."""
    warnings.simplefilter("ignore")
    x = _clip_with_silence()
    y_ref, det = oracle.sim(x, FS, return_details=True)
    assert np.isnan(y_ref).any()
    y, lists = repet._host.sim_f64(x, FS, repet._tunables(), return_indices=True)
    assert [len(v) for v in lists] == [len(v) for v in det["indices"]]
    assert all(np.array_equal(a, b) for a, b in zip(lists, det["indices"]))
    _assert_same(y, y_ref, "sim with silence")


def test_silent_stretch_simonline(repet):
    """The same for the online method: silence inside the 10 s buffer (NaN similarities block their neighbours)
    and silent target frames (empty lists, NaN output over those frames)."""
    warnings.simplefilter("ignore")
    x = repet_synth.make_clip(78, 14 * FS).T.astype(np.float64)
    x[int(6.2 * FS) : int(6.9 * FS)] = 0.0
    x[int(11.0 * FS) : int(11.8 * FS)] = 0.0
    y_ref, det = oracle.simonline(x, FS, return_details=True)
    y, lists = repet._host.simonline_f64(x, FS, repet._tunables(), return_indices=True)
    first = det["first_frame"]
    assert [len(v) for v in lists[first:]] == [len(v) for v in det["indices"]]
    assert all(np.array_equal(a, b) for a, b in zip(lists[first:], det["indices"]))
    _assert_same(y, y_ref, "simonline with silence")


def test_all_silent_clip(repet):
    warnings.simplefilter("ignore")
    z = np.zeros((8 * FS, 2))
    y_ref, det = oracle.original(z, FS, return_details=True)
    y, period = repet._host.original_f64(z, FS, repet._tunables(), return_period=True)
    assert period == det["period"]  # argmax of an all-zero beat spectrum: the first lag
    _assert_same(y, y_ref, "silent original")
    y_ref, det = oracle.sim(z, FS, return_details=True)
    y, lists = repet._host.sim_f64(z, FS, repet._tunables(), return_indices=True)
    assert [len(v) for v in lists] == [len(v) for v in det["indices"]]
    _assert_same(y, y_ref, "silent sim")


def test_empty_batch_and_odd_lengths(repet):
    empty = np.zeros((0, 2, 6 * FS), dtype=np.float32)
    background, periods = repet.original_batch(empty, FS)
    assert background.shape == empty.shape and periods.shape == (0,)
    for samples in (4 * FS + 1, 4 * FS + 1023, 4 * FS + 1025):  # around the hop: frame count changes by one
        x = repet_synth.make_clip(5, samples).T.astype(np.float64)
        y_ref, det = oracle.original(x, FS, return_details=True)
        y, period = repet._host.original_f64(x, FS, repet._tunables(), return_period=True)
        assert period == det["period"] and y.shape == x.shape
        _assert_same(y, y_ref, "odd length %d" % samples)


def test_clip_shorter_than_the_period_range_raises(repet):
    with pytest.raises(ValueError):  # the reference: argmax of an empty sequence (quirk Q17)
        repet.original(np.full((3000, 2), 0.01), FS)
    with pytest.raises(ValueError):
        repet.original_batch(np.zeros((2, 2, 3000), dtype=np.float32), FS)


def test_ragged_batch_equals_single_calls(repet):
    """Clips of different lengths in one call: grouped by shape, never padded (a clip's frame count and period
    range depend on its length), outputs in input order."""
    lengths = [7 * FS, 9 * FS + 11, 7 * FS, 5 * FS + 3, 9 * FS + 11]
    clips = [repet_synth.make_clip(700 + i, n) for i, n in enumerate(lengths)]
    backgrounds, periods = repet.original_batch(clips, FS)
    assert len(backgrounds) == len(clips) and periods.shape == (len(clips),)
    for i, clip in enumerate(clips):
        single, period = repet.original_batch(clip[None], FS)
        assert backgrounds[i].shape == clip.shape
        assert np.array_equal(backgrounds[i], single[0]) and int(periods[i]) == int(period[0])
        _, det = oracle.original(clip.T.astype(np.float64), FS, return_details=True)
        assert int(periods[i]) == det["period"]
