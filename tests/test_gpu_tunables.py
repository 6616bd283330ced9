"""GPU: the nine module tunables (repet.py:42-63) varied together, every driver against the oracle -- on the fast path
(44.1 kHz stereo) and on the general path (3 channels), integers bit-exact, signals within tolerance.  The reference
reads the tunables at call time; so does the drop-in module."""

import numpy as np
import pytest

import repet_oracle as oracle
import repet_synth

pytestmark = pytest.mark.gpu

FS = 44100

# (name, tunables); chosen to reach the code paths the defaults do not: list lengths above 32 and above 128 (the
# shared-memory median), distance 0, a positive threshold, even / large filter orders, short and long segments,
# narrow period ranges, no high-pass
CASES = [
    ("wide_lists", dict(similarity_number=160, similarity_distance=0.05, buffer_length=12)),
    ("no_distance", dict(similarity_distance=0, similarity_number=40, similarity_threshold=0.9)),
    ("threshold", dict(similarity_threshold=0.97, similarity_number=7)),
    ("order_even_large", dict(filter_order=8, segment_length=6, segment_step=2)),
    ("order_17", dict(filter_order=17, segment_length=8, segment_step=4)),
    ("narrow_period", dict(period_range=[0.5, 2.5], cutoff_frequency=0)),
    ("short_segments", dict(segment_length=4, segment_step=1, period_range=[0.3, 3], cutoff_frequency=300)),
    ("tiny_buffer", dict(buffer_length=2.5, similarity_number=3, similarity_distance=0.2)),
]


@pytest.fixture(scope="module")
def repet():
    import repet as module

    module._host.get_handle(0)
    return module


def _check(repet, x, fs, overrides, tol, what):
    saved = {name: getattr(repet, name) for name in overrides}
    try:
        for name, value in overrides.items():
            setattr(repet, name, value)
        tun = repet._tunables()
        got = {
            "original": repet._host.original_f64(x, fs, tun, return_period=True),
            "extended": repet._host.extended_f64(x, fs, tun, return_periods=True),
            "adaptive": repet._host.adaptive_f64(x, fs, tun, return_periods=True),
            "sim": repet._host.sim_f64(x, fs, tun, return_indices=True),
            "simonline": repet._host.simonline_f64(x, fs, tun, return_indices=True),
        }
    finally:
        for name, value in saved.items():
            setattr(repet, name, value)
    for method, (y, ints) in got.items():
        with np.errstate(all="ignore"):
            y_ref, det = getattr(oracle, method)(x, fs, return_details=True, **overrides)
        label = "%s / %s" % (what, method)
        if method == "original":
            assert ints == det["period"], label
        elif method in ("extended", "adaptive"):
            assert np.array_equal(np.asarray(ints), np.asarray(det["periods"])), label
        else:
            first = det.get("first_frame", 0)
            assert len(ints) - first == len(det["indices"]), label
            bad = [i for i, (a, b) in enumerate(zip(ints[first:], det["indices"])) if not np.array_equal(a, b)]
            assert not bad, "%s: lists differ at frames %s" % (label, bad[:6])
        assert y.shape == y_ref.shape and np.array_equal(np.isnan(y), np.isnan(y_ref)), label
        peak = float(np.nanmax(np.abs(y_ref)))
        err = float(np.nanmax(np.abs(y - y_ref))) / peak
        assert err <= tol, "%s: max-abs/peak %.3e" % (label, err)


@pytest.mark.parametrize("name,overrides", CASES)
def test_tunables_fast_path(repet, name, overrides):
    x = repet_synth.make_clip(1200 + len(name), 15 * FS + 511).T.astype(np.float64)
    _check(repet, x, FS, overrides, 1e-4, "fast " + name)


@pytest.mark.parametrize("name,overrides", CASES[:4] + CASES[5:6])
def test_tunables_general_path(repet, name, overrides):
    x = repet_synth.make_clip(1300 + len(name), 13 * FS + 77, 3).T.astype(np.float64)
    _check(repet, x, FS, overrides, 1e-9, "general " + name)
