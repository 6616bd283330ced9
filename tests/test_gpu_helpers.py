"""GPU: the remaining private helpers of repet.py through the C ABI, against the golden helper
vectors recorded from the reference (tests/golden/helpers.npz) and the oracle."""

import numpy as np
import pytest

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


def test_acorr_matches_reference(repet, golden_helpers):
    V = make_golden.helper_inputs()["spectrogram"]
    got = repet._acorr(V.T)
    ref = golden_helpers["acorr"]
    assert got.shape == ref.shape
    # fp32 transforms: the error lives in the raw correlation sums, so compare those (the unbiased
    # 1/(rows - lag) factor amplifies it arbitrarily at the last lags)
    weight = np.arange(V.shape[1], 0, -1)[:, None]
    assert float(np.max(np.abs(got - ref) * weight) / np.max(np.abs(ref) * weight)) <= 2e-6
    assert float(np.max(np.abs(got[:150] - ref[:150])) / np.max(np.abs(ref))) <= 1e-5


def test_similarity_helpers_are_exact(repet, golden_helpers):
    V = make_golden.helper_inputs()["spectrogram"]
    Vq = V.astype(np.float32).astype(np.float64)  # the ABI takes fp32 magnitudes
    S = repet._selfsimilaritymatrix(V)
    assert float(np.max(np.abs(S - oracle.selfsimilaritymatrix(Vq)))) <= 1e-14
    assert float(np.max(np.abs(S - golden_helpers["selfsim"]))) <= 1e-6
    s = repet._similaritymatrix(V, V[:, 10:11])
    assert s.shape == golden_helpers["sim"].shape
    assert float(np.max(np.abs(s - oracle.similaritymatrix(Vq, Vq[:, 10:11])))) <= 1e-14


def test_localmaxima_and_indices_bit_exact(repet, golden_helpers):
    h = make_golden.helper_inputs()
    values, index = repet._localmaxima(h["vector"], 0.2, 7, 20)
    assert np.array_equal(index, golden_helpers["localmaxima_indices"])
    assert np.array_equal(values, golden_helpers["localmaxima_values"])
    # on the REFERENCE's float64 similarity matrix: pure comparisons, must be bit-exact
    lists = repet._indices(golden_helpers["selfsim"], 0, 9, 12)
    assert np.array_equal(np.array([len(v) for v in lists]), golden_helpers["indices_counts"])
    assert np.array_equal(np.concatenate(lists), golden_helpers["indices_flat"])
    # NaN never wins and blocks its neighbours; plateaus give no maximum (quirk Q7)
    v = np.array([0.1, 0.5, np.nan, 0.9, 0.2, 0.7, 0.7, 0.1, 0.8])
    got = repet._localmaxima(v, 0.0, 1, 10)[1]
    assert np.array_equal(got, oracle.localmaxima(v, 0.0, 1, 10)[1])


def test_simmask_matches_oracle(repet):
    x = repet_synth.make_clip(17, int(7.0 * FS)).astype(np.float64)
    N, w, H = oracle.stft_parameters(FS)
    V = np.abs(oracle.stft(x[0], w, H)[: N // 2 + 1])
    Vq = V.astype(np.float32).astype(np.float64)
    S = oracle.selfsimilaritymatrix(V)
    for distance, number in ((43, 100), (3, 40)):  # short lists (networks) and long lists (quickselect)
        lists = oracle.indices(S, 0, distance, number)
        M = repet._simmask(V, lists)
        assert float(np.max(np.abs(M - oracle.simmask(Vq, lists)))) <= 2e-6, (distance, number)
    assert max(len(v) for v in oracle.indices(S, 0, 3, 40)) > 32
