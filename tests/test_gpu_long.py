"""GPU: parity at the sizes of BASELINE.json configs 3-5 against goldens recorded from the UNMODIFIED reference
(oracle/make_golden_long.py -> tests/golden/long_*.npz): a 10-minute `adaptive` track (25 841 per-frame periods),
one hour of `extended` (719 segment periods) and of `simonline` (155 039 frames of similar-frame lists), and 5- and
10-minute `sim` tracks (12 921 / 25 841 lists, float64 samples that fp32 cannot hold).  Integers bit-exact (lists:
lengths in full, contents and order through one digest per block of 64 lists); signals within 1e-4 of the
reference's decimated samples, rms and peak.  The inputs are regenerated from their seeds (one CPU core, up to
~40 s for an hour of audio)."""

import os
import warnings

import numpy as np
import pytest

import make_golden
import make_golden_long

pytestmark = pytest.mark.gpu

FS = 44100
RTOL_SIGNAL = 1e-4
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def repet():
    import repet as module

    module._host.get_handle(0)
    return module


def _golden(case):
    path = os.path.join(GOLDEN, "long_%s.npz" % case)
    if not os.path.isfile(path):
        pytest.skip("golden %s not generated (oracle/make_golden_long.py --case %s)" % (path, case))
    return np.load(path)


def _assert_decimated(y, golden, what):
    assert y.shape[0] == int(golden["samples"]), what
    assert np.all(np.isfinite(y)), what
    dec = y[:: int(golden["decimate"])]
    ref = golden["dec"]
    assert dec.shape == ref.shape, what
    peak = float(golden["max"])
    rel = float(np.linalg.norm(np.ravel(dec - ref)) / np.linalg.norm(np.ravel(ref)))
    worst = float(np.max(np.abs(dec - ref))) / peak
    assert rel <= RTOL_SIGNAL and worst <= RTOL_SIGNAL, "%s: rel L2 %.3e, max-abs/peak %.3e" % (what, rel, worst)
    assert abs(float(np.sqrt(np.mean(np.square(y)))) - float(golden["rms"])) <= RTOL_SIGNAL * float(golden["rms"]), what
    assert abs(float(np.max(np.abs(y))) - peak) <= RTOL_SIGNAL * peak, what


def _assert_lists(lists, golden, what):
    counts = np.array([len(v) for v in lists], dtype=np.int64)
    assert counts.shape == golden["counts"].shape, what
    bad = np.nonzero(counts != golden["counts"].astype(np.int64))[0]
    assert bad.size == 0, "%s: %d list lengths differ, first at frame %d" % (what, bad.size, bad[0])
    assert int(counts.sum()) == int(golden["total"]), what
    digests = make_golden.list_digests(lists, make_golden_long.BLOCK)
    bad = np.nonzero(digests != golden["digests"])[0]
    assert bad.size == 0, "%s: lists differ in blocks of %d frames starting at %s" % (
        what, make_golden_long.BLOCK, (bad[:8] * make_golden_long.BLOCK).tolist())


@pytest.mark.parametrize("case", ["adaptive_1min", "adaptive_10min"])
def test_adaptive_long_track(repet, case):
    """BASELINE configs[2] size: every per-frame period of a 10-minute track (repet.py:483-568)."""
    golden = _golden(case)
    x = make_golden_long.long_input(case)
    y, periods = repet._host.adaptive_f64(x, FS, repet._tunables(), return_periods=True)
    ref = golden["periods"].astype(np.int64)
    assert periods.shape == ref.shape
    bad = np.nonzero(periods != ref)[0]
    assert bad.size == 0, "%d of %d per-frame periods differ, first at frame %d" % (bad.size, ref.size, bad[0] if bad.size else -1)
    _assert_decimated(y, golden, case)


def test_extended_one_hour(repet):
    """BASELINE configs[4] size: the periods of all 719 segments of one hour (repet.py:263-419)."""
    golden = _golden("extended_1h")
    x = make_golden_long.long_input("extended_1h")
    y, periods = repet._host.extended_f64(x, FS, repet._tunables(), return_periods=True)
    ref = golden["periods"].astype(np.int64)
    assert periods.shape == ref.shape == (719,)
    bad = np.nonzero(periods != ref)[0]
    assert bad.size == 0, "segment periods differ at %s" % bad[:8].tolist()
    _assert_decimated(y, golden, "extended_1h")


@pytest.mark.parametrize("case", ["sim_5min", "sim_10min"])
def test_sim_long_tracks(repet, case):
    """BASELINE configs[3] size: the similar-frame lists (set and order) of 12 921 / 25 841 frames (repet.py:631-709)."""
    warnings.simplefilter("ignore")
    golden = _golden(case)
    x = make_golden_long.long_input(case)
    y, lists = repet._host.sim_f64(x, FS, repet._tunables(), return_indices=True)
    _assert_lists(lists, golden, case)
    _assert_decimated(y, golden, case)


@pytest.mark.parametrize("case", ["simonline_1min", "simonline_1h"])
def test_simonline_long_stream(repet, case):
    """BASELINE configs[4] size: one hour through the online REPET-SIM, S = 155 038 * 1024 + 2048 (repet.py:771-911);
    lists of frames >= buffer_frames - 1 in ring-slot order (quirk Q6), nothing before them (quirk Q5)."""
    golden = _golden(case)
    x = make_golden_long.long_input(case)
    y, lists = repet._host.simonline_f64(x, FS, repet._tunables(), return_indices=True)
    first = int(golden["first_frame"])
    assert all(len(v) == 0 for v in lists[:first])
    _assert_lists(lists[first:], golden, case)
    assert np.all(y[: first * 1024] == 0)
    _assert_decimated(y, golden, case)
