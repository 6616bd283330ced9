"""
CPU oracle for the REPET separation hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a float64 NumPy restatement of the algorithm in the reference
`repet.py` (zafarrafii/REPET-Python).  It exists so that the CUDA path can be checked
against an independent CPU statement of the same arithmetic on a box where the
reference itself is not present.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product
(`repet-python_b200/`) never does: it fails loudly when its CUDA library is missing.

Parity status: PINNED.  `oracle/make_golden.py` imports the unmodified reference from
`/root/reference/repet.py` (through the 6-line shim of SURVEY.md section 8(c)) and
(1) asserts that every function below reproduces the reference on the bundled
`audio_file.wav` excerpt and on seeded synthetic clips -- integer outputs bit-exact,
floats to <= 1e-12 relative -- and (2) writes the golden vectors under `tests/golden/`
that `tests/test_oracle_golden.py` re-checks on every run.  The reference has no tests
or golden vectors of its own.

Third-party arithmetic the reference delegates to (not vendored, not pinned by the
reference, README.md:22): NumPy pocketfft / BLAS matmul / median / argmax / argsort and
the SciPy window constructors.  The oracle calls the same NumPy/SciPy entry points with
the same argument shapes wherever a result could depend on them.

Every function cites the reference lines (`repet.py:a-b`) it restates.  Where the
reference loops in Python the oracle vectorises, keeping the floating-point operations
per output element the same.
"""

import numpy as np
import scipy.signal.windows

EPS = np.finfo(float).eps  # repet.py:1446, 1504, 1541, 882

# Tunables of the reference module (repet.py:42-63).  The oracle entry points take them
# as keyword arguments so tests can vary them without mutating module state.
DEFAULTS = dict(
    cutoff_frequency=100,
    period_range=(1, 10),
    segment_length=10,
    segment_step=5,
    filter_order=5,
    similarity_threshold=0,
    similarity_distance=1,
    similarity_number=100,
    buffer_length=10,
)


def _cfg(overrides):
    cfg = dict(DEFAULTS)
    for key, value in overrides.items():
        if key not in cfg:
            raise TypeError("unknown tunable %r" % key)
        cfg[key] = value
    return cfg


# --------------------------------------------------------------------------------------
# derived parameters
# --------------------------------------------------------------------------------------
def stft_parameters(sampling_frequency):
    """Window length N, periodic Hamming window, step H (repet.py:130-132, 289-291,
    488-490, 636-638, 776-778)."""
    window_length = pow(2, int(np.ceil(np.log2(0.04 * sampling_frequency))))
    window_function = scipy.signal.windows.hamming(window_length, sym=False)
    step_length = int(window_length / 2)
    return window_length, window_function, step_length


def number_of_frames(number_samples, window_length, step_length):
    """Frame count of the centred STFT (repet.py:135-146, 1018-1028)."""
    padding_length = int(np.floor(window_length / 2))
    return int(np.ceil(((number_samples + 2 * padding_length) - window_length) / step_length)) + 1


def period_range_frames(period_range, sampling_frequency, step_length):
    """Period range in frames, NumPy round-half-even (repet.py:165-167)."""
    return np.round(np.array(period_range) * sampling_frequency / step_length).astype(int)


def cutoff_bins(cutoff_frequency, window_length, sampling_frequency):
    """High-pass cutoff in bins, Python round-half-even (repet.py:173)."""
    return round(cutoff_frequency * window_length / sampling_frequency)


# --------------------------------------------------------------------------------------
# L1 helpers (repet.py:1001-1545)
# --------------------------------------------------------------------------------------
def stft(audio_signal, window_function, step_length):
    """Centred, zero-padded, windowed full complex FFT per frame (repet.py:1001-1060).
    Returns (window_length, number_times) complex128."""
    audio_signal = np.asarray(audio_signal, dtype=float)
    number_samples = len(audio_signal)
    window_length = len(window_function)
    padding_length = int(np.floor(window_length / 2))
    number_times = number_of_frames(number_samples, window_length, step_length)
    total = number_times * step_length + (window_length - step_length)
    padded = np.zeros(total)
    padded[padding_length : padding_length + number_samples] = audio_signal
    # frame j = padded[j*H : j*H+N] * w  (repet.py:1051-1055), as one strided gather
    frame_index = np.arange(window_length)[:, None] + step_length * np.arange(number_times)[None, :]
    frames = padded[frame_index] * window_function[:, None]
    return np.fft.fft(frames, axis=0)  # repet.py:1058


def istft(audio_stft, window_function, step_length):
    """real(ifft) per frame, overlap-add, trim, divide by the COLA gain
    (repet.py:1063-1105).  No synthesis window."""
    window_length, number_times = np.shape(audio_stft)
    number_samples = number_times * step_length + (window_length - step_length)
    frames = np.real(np.fft.ifft(audio_stft, axis=0))  # repet.py:1085
    audio_signal = np.zeros(number_samples)
    # With N = 2H every sample receives at most two frames; add them in frame order as the
    # reference loop does (repet.py:1089-1095).  Even frames never overlap each other,
    # nor do odd frames, so two strided adds reproduce the same sums.
    if window_length == 2 * step_length:
        for parity in (0, 1):
            sel = frames[:, parity::2]
            count = sel.shape[1]
            if count == 0:
                continue
            start = parity * step_length
            view = audio_signal[start : start + count * window_length].reshape(count, window_length)
            view += sel.T
    else:
        i = 0
        for j in range(number_times):
            audio_signal[i : i + window_length] += frames[:, j]
            i += step_length
    audio_signal = audio_signal[window_length - step_length : number_samples - (window_length - step_length)]
    return audio_signal / sum(window_function[0:window_length:step_length])  # repet.py:1103


def acorr(data_matrix):
    """Unbiased autocorrelation of every column by Wiener-Khinchin with FFT length
    2*rows (repet.py:1108-1139)."""
    number_rows = data_matrix.shape[0]
    psd = np.power(np.abs(np.fft.fft(data_matrix, n=2 * number_rows, axis=0)), 2)
    autocorrelation = np.real(np.fft.ifft(psd, axis=0))[0:number_rows, :]
    return np.divide(autocorrelation, np.arange(number_rows, 0, -1)[:, np.newaxis])


def beatspectrum(audio_spectrogram):
    """Mean over frequency of the row-wise autocorrelation (repet.py:1142-1158)."""
    return np.mean(acorr(audio_spectrogram.T), axis=1)


def beatspectrogram(audio_spectrogram, segment_length, segment_step):
    """Centred sliding beat spectrum, replicated over step-1 columns -- column
    i+step-1 stays zero (quirk Q3) (repet.py:1161-1206)."""
    number_times = np.shape(audio_spectrogram)[1]
    left = int(np.ceil((segment_length - 1) / 2))
    right = int(np.floor((segment_length - 1) / 2))
    padded = np.pad(audio_spectrogram, ((0, 0), (left, right)), "constant", constant_values=0)
    beat_spectrogram = np.zeros((segment_length, number_times))
    for i in range(0, number_times, segment_step):
        column = beatspectrum(padded[:, i : i + segment_length])
        beat_spectrogram[:, i] = column
        stop = min(i + segment_step - 1, number_times)
        if stop > i:
            beat_spectrogram[:, i:stop] = column[:, np.newaxis]
    return beat_spectrogram


def periods(beat_spectrogram, period_range):
    """argmax over lags [lo, min(hi, n_lags//3)) plus 1 plus lo (quirks Q1, Q2)
    (repet.py:1249-1291).  1-D input -> scalar, 2-D input -> one period per column."""
    lo = int(period_range[0])
    hi = min(int(period_range[1]), int(np.floor(beat_spectrogram.shape[0] / 3)))
    if beat_spectrogram.ndim == 1:
        return np.argmax(beat_spectrogram[lo:hi]) + 1 + lo
    return np.argmax(beat_spectrogram[lo:hi, :], axis=0) + 1 + lo


def selfsimilaritymatrix(data_matrix):
    """Column-normalised Gram matrix (repet.py:1209-1225)."""
    data_matrix = data_matrix / np.sqrt(np.sum(np.power(data_matrix, 2), axis=0))
    return np.matmul(data_matrix.T, data_matrix)


def similaritymatrix(data_matrix1, data_matrix2):
    """Cosine similarity between the columns of two matrices (repet.py:1228-1246)."""
    data_matrix1 = data_matrix1 / np.sqrt(np.sum(np.power(data_matrix1, 2), axis=0))
    data_matrix2 = data_matrix2 / np.sqrt(np.sum(np.power(data_matrix2, 2), axis=0))
    return np.matmul(data_matrix1.T, data_matrix2)


def localmaxima(data_vector, minimum_value, minimum_distance, number_values):
    """Strict local maxima within +-distance (windows clipped at the ends), >= threshold,
    best `number_values` by value descending (quirks Q7, Q8) (repet.py:1294-1345).
    The element-by-element scan of the reference becomes two sliding-window maxima."""
    data_vector = np.asarray(data_vector, dtype=float)
    number_elements = len(data_vector)
    d = int(minimum_distance)
    if number_elements == 0:
        empty = np.array([], dtype=int)
        return data_vector[empty], empty
    if d > 0:
        padded = np.concatenate((np.full(d, -np.inf), data_vector, np.full(d, -np.inf)))
        windows = np.lib.stride_tricks.sliding_window_view(padded, d)
        left_max = windows[:number_elements].max(axis=1)  # elements i-d .. i-1
        right_max = windows[d + 1 : d + 1 + number_elements].max(axis=1)  # i+1 .. i+d
        with np.errstate(invalid="ignore"):
            keep = (data_vector >= minimum_value) & (data_vector > left_max) & (data_vector > right_max)
    else:
        with np.errstate(invalid="ignore"):
            keep = data_vector >= minimum_value
    maximum_indices = np.flatnonzero(keep)
    maximum_values = data_vector[maximum_indices]
    sort_indices = np.argsort(maximum_values)[::-1]  # repet.py:1335 (same call, same tie order)
    sort_indices = sort_indices[0 : min(number_values, len(maximum_values))]
    return maximum_values[sort_indices], maximum_indices[sort_indices]


def indices(similarity_matrix, similarity_threshold, similarity_distance, similarity_number):
    """localmaxima of every column (repet.py:1348-1383).  Ragged list of int arrays."""
    number_times = similarity_matrix.shape[0]
    return [
        localmaxima(similarity_matrix[:, i], similarity_threshold, similarity_distance, similarity_number)[1]
        for i in range(number_times)
    ]


def _softmask(audio_spectrogram, repeating_spectrogram):
    """min with the mixture, then (W+eps)/(V+eps) (quirk Q10) (repet.py:1441-1448,
    1501-1506, 1538-1543)."""
    repeating_spectrogram = np.minimum(audio_spectrogram, repeating_spectrogram)
    return (repeating_spectrogram + EPS) / (audio_spectrogram + EPS)


def mask(audio_spectrogram, repeating_period):
    """Median over the period-strided frames of each phase; the zero padding of the last
    segment is excluded (quirk Q9) (repet.py:1386-1458)."""
    number_frequencies, number_times = np.shape(audio_spectrogram)
    p = int(repeating_period)
    model = np.empty((number_frequencies, p))
    with np.errstate(invalid="ignore"):
        for q in range(p):
            # columns q, q+p, ... < T: r of them for q < T-(r-1)p, r-1 otherwise
            model[:, q] = np.median(audio_spectrogram[:, q::p], axis=1) if q < number_times else np.nan
    return _softmask(audio_spectrogram, model[:, np.arange(number_times) % p])


def adaptivemask(audio_spectrogram, repeating_periods, filter_order):
    """Per-frame median over the in-range frames i + c*p_i (repet.py:1461-1508)."""
    number_frequencies, number_times = np.shape(audio_spectrogram)
    center_indices = np.arange(1, filter_order + 1) - int(np.ceil(filter_order / 2))
    repeating_spectrogram = np.zeros((number_frequencies, number_times))
    for i in range(number_times):
        all_indices = i + center_indices * repeating_periods[i]
        all_indices = all_indices[np.logical_and(all_indices >= 0, all_indices < number_times)]
        repeating_spectrogram[:, i] = np.median(audio_spectrogram[:, all_indices], axis=1)
    return _softmask(audio_spectrogram, repeating_spectrogram)


def simmask(audio_spectrogram, similarity_indices):
    """Per-frame median over the listed similar frames (repet.py:1511-1545)."""
    number_frequencies, number_times = np.shape(audio_spectrogram)
    repeating_spectrogram = np.zeros((number_frequencies, number_times))
    with np.errstate(invalid="ignore"):
        for i in range(number_times):
            repeating_spectrogram[:, i] = np.median(audio_spectrogram[:, similarity_indices[i]], 1)
    return _softmask(audio_spectrogram, repeating_spectrogram)


# --------------------------------------------------------------------------------------
# shared driver pieces
# --------------------------------------------------------------------------------------
def _analysis(audio_signal, window_function, step_length):
    """STFT of every channel + magnitude half-spectrogram (repet.py:149-158)."""
    number_samples, number_channels = np.shape(audio_signal)
    window_length = len(window_function)
    number_times = number_of_frames(number_samples, window_length, step_length)
    audio_stft = np.zeros((window_length, number_times, number_channels), dtype=complex)
    for i in range(number_channels):
        audio_stft[:, :, i] = stft(audio_signal[:, i], window_function, step_length)
    audio_spectrogram = abs(audio_stft[0 : int(window_length / 2) + 1, :, :])
    return audio_stft, audio_spectrogram


def _synthesis(repeating_mask, audio_stft_channel, window_function, step_length, cutoff2, number_samples):
    """High-pass rows 1..cutoff2 (quirk Q11), mirror, apply, ISTFT, truncate
    (repet.py:185-200)."""
    repeating_mask = np.array(repeating_mask)
    repeating_mask[1 : cutoff2 + 1, :] = 1
    repeating_mask = np.concatenate((repeating_mask, repeating_mask[-2:0:-1, :]), axis=0)
    return istft(repeating_mask * audio_stft_channel, window_function, step_length)[0:number_samples]


# --------------------------------------------------------------------------------------
# L2 drivers (repet.py:67-911)
# --------------------------------------------------------------------------------------
def original(audio_signal, sampling_frequency, return_details=False, **tunables):
    """repet.original (repet.py:125-202)."""
    cfg = _cfg(tunables)
    audio_signal = np.asarray(audio_signal, dtype=float)
    number_samples, number_channels = np.shape(audio_signal)
    window_length, window_function, step_length = stft_parameters(sampling_frequency)
    audio_stft, audio_spectrogram = _analysis(audio_signal, window_function, step_length)
    beat_spectrum = beatspectrum(np.power(np.mean(audio_spectrogram, axis=2), 2))
    period_range2 = period_range_frames(cfg["period_range"], sampling_frequency, step_length)
    repeating_period = periods(beat_spectrum, period_range2)
    cutoff2 = cutoff_bins(cfg["cutoff_frequency"], window_length, sampling_frequency)
    background_signal = np.zeros((number_samples, number_channels))
    for i in range(number_channels):
        repeating_mask = mask(audio_spectrogram[:, :, i], repeating_period)
        background_signal[:, i] = _synthesis(
            repeating_mask, audio_stft[:, :, i], window_function, step_length, cutoff2, number_samples
        )
    if return_details:
        return background_signal, dict(period=int(repeating_period), beat_spectrum=beat_spectrum)
    return background_signal


def extended(audio_signal, sampling_frequency, return_details=False, **tunables):
    """repet.extended: `original` per 10 s segment every 5 s, triangular cross-fade
    (quirk Q15) (repet.py:263-419)."""
    cfg = _cfg(tunables)
    audio_signal = np.asarray(audio_signal, dtype=float)
    number_samples, number_channels = np.shape(audio_signal)
    segment_length2 = round(cfg["segment_length"] * sampling_frequency)
    segment_step2 = round(cfg["segment_step"] * sampling_frequency)
    segment_overlap2 = segment_length2 - segment_step2
    if number_samples < segment_length2 + segment_step2:
        number_segments = 1
    else:
        number_segments = 1 + int(np.floor((number_samples - segment_length2) / segment_step2))
        segment_window = scipy.signal.windows.triang(2 * segment_overlap2)
    window_length, window_function, step_length = stft_parameters(sampling_frequency)
    period_range2 = period_range_frames(cfg["period_range"], sampling_frequency, step_length)
    cutoff2 = cutoff_bins(cfg["cutoff_frequency"], window_length, sampling_frequency)
    background_signal = np.zeros((number_samples, number_channels))
    segment_periods = []
    k = 0
    for j in range(number_segments):
        if number_segments == 1:
            audio_segment = audio_signal
            segment_length2 = number_samples
        elif j < number_segments - 1:
            audio_segment = audio_signal[k : k + segment_length2, :]
        else:
            audio_segment = audio_signal[k:number_samples, :]
            segment_length2 = len(audio_segment)
        audio_stft, audio_spectrogram = _analysis(audio_segment, window_function, step_length)
        beat_spectrum = beatspectrum(np.power(np.mean(audio_spectrogram, axis=2), 2))
        repeating_period = periods(beat_spectrum, period_range2)
        segment_periods.append(int(repeating_period))
        background_segment = np.zeros((segment_length2, number_channels))
        for i in range(number_channels):
            repeating_mask = mask(audio_spectrogram[:, :, i], repeating_period)
            background_segment[:, i] = _synthesis(
                repeating_mask, audio_stft[:, :, i], window_function, step_length, cutoff2, segment_length2
            )
        if number_segments == 1:
            background_signal = background_segment
        else:
            if j == 0:
                background_signal[0:segment_length2, :] += background_segment
            else:
                background_signal[k : k + segment_overlap2, :] *= segment_window[
                    segment_overlap2 : 2 * segment_overlap2, np.newaxis
                ]
                background_segment[0:segment_overlap2, :] *= segment_window[0:segment_overlap2, np.newaxis]
                background_signal[k : k + segment_length2, :] += background_segment
            k = k + segment_step2
    if return_details:
        return background_signal, dict(periods=segment_periods)
    return background_signal


def adaptive(audio_signal, sampling_frequency, return_details=False, **tunables):
    """repet.adaptive (repet.py:483-568)."""
    cfg = _cfg(tunables)
    audio_signal = np.asarray(audio_signal, dtype=float)
    number_samples, number_channels = np.shape(audio_signal)
    window_length, window_function, step_length = stft_parameters(sampling_frequency)
    audio_stft, audio_spectrogram = _analysis(audio_signal, window_function, step_length)
    segment_length2 = int(round(cfg["segment_length"] * sampling_frequency / step_length))
    segment_step2 = int(round(cfg["segment_step"] * sampling_frequency / step_length))
    beat_spectrogram = beatspectrogram(
        np.power(np.mean(audio_spectrogram, axis=2), 2), segment_length2, segment_step2
    )
    period_range2 = period_range_frames(cfg["period_range"], sampling_frequency, step_length)
    repeating_periods = periods(beat_spectrogram, period_range2)
    cutoff2 = cutoff_bins(cfg["cutoff_frequency"], window_length, sampling_frequency)
    background_signal = np.zeros((number_samples, number_channels))
    for i in range(number_channels):
        repeating_mask = adaptivemask(audio_spectrogram[:, :, i], repeating_periods, cfg["filter_order"])
        background_signal[:, i] = _synthesis(
            repeating_mask, audio_stft[:, :, i], window_function, step_length, cutoff2, number_samples
        )
    if return_details:
        return background_signal, dict(periods=np.asarray(repeating_periods), beat_spectrogram=beat_spectrogram)
    return background_signal


def sim(audio_signal, sampling_frequency, return_details=False, **tunables):
    """repet.sim (repet.py:631-709)."""
    cfg = _cfg(tunables)
    audio_signal = np.asarray(audio_signal, dtype=float)
    number_samples, number_channels = np.shape(audio_signal)
    window_length, window_function, step_length = stft_parameters(sampling_frequency)
    audio_stft, audio_spectrogram = _analysis(audio_signal, window_function, step_length)
    similarity_matrix = selfsimilaritymatrix(np.mean(audio_spectrogram, axis=2))
    similarity_distance2 = int(round(cfg["similarity_distance"] * sampling_frequency / step_length))
    similarity_indices = indices(
        similarity_matrix, cfg["similarity_threshold"], similarity_distance2, cfg["similarity_number"]
    )
    cutoff2 = cutoff_bins(cfg["cutoff_frequency"], window_length, sampling_frequency)
    background_signal = np.zeros((number_samples, number_channels))
    for i in range(number_channels):
        repeating_mask = simmask(audio_spectrogram[:, :, i], similarity_indices)
        background_signal[:, i] = _synthesis(
            repeating_mask, audio_stft[:, :, i], window_function, step_length, cutoff2, number_samples
        )
    if return_details:
        return background_signal, dict(indices=similarity_indices, similarity_matrix=similarity_matrix)
    return background_signal


def simonline(audio_signal, sampling_frequency, return_details=False, **tunables):
    """repet.simonline (repet.py:771-911), restated frame-parallel: every frame's result
    depends only on the magnitudes of the previous buffer_frames-1 input frames, visited
    in ring-buffer SLOT order (quirk Q6).  No centring pad; frames before
    buffer_frames-1 are never synthesised (quirk Q5)."""
    cfg = _cfg(tunables)
    audio_signal = np.asarray(audio_signal, dtype=float)
    number_samples, number_channels = np.shape(audio_signal)
    window_length, window_function, step_length = stft_parameters(sampling_frequency)
    number_times = int(np.ceil((number_samples - window_length) / step_length + 1))
    number_frequencies = int(window_length / 2 + 1)
    buffer_length2 = round((cfg["buffer_length"] * sampling_frequency) / step_length)
    # the warm-up loop multiplies a truncated slice by the window when the signal is too
    # short and NumPy raises (repet.py:801-804, quirk Q5 / Q17)
    if (buffer_length2 - 2) * step_length + window_length > number_samples:
        raise ValueError("operands could not be broadcast together (signal shorter than the buffer)")
    total = (number_times - 1) * step_length + window_length
    padded = np.zeros((total, number_channels))
    padded[0:number_samples, :] = audio_signal
    similarity_distance2 = int(round(cfg["similarity_distance"] * sampling_frequency / step_length))
    cutoff2 = cutoff_bins(cfg["cutoff_frequency"], window_length, sampling_frequency)
    # all frame spectra up front (repet.py:801, 846)
    frame_index = np.arange(window_length)[:, None] + step_length * np.arange(number_times)[None, :]
    frame_ft = np.empty((window_length, number_times, number_channels), dtype=complex)
    for i in range(number_channels):
        frame_ft[:, :, i] = np.fft.fft(padded[:, i][frame_index] * window_function[:, None], axis=0)
    magnitude = np.abs(frame_ft[0:number_frequencies, :, :])
    background_signal = np.zeros((total, number_channels))
    all_indices = []
    slots = np.arange(buffer_length2)
    for j in range(buffer_length2 - 1, number_times):
        j0 = j % buffer_length2
        # slot b holds frame j-(j0-b) if b <= j0 else j-(j0-b)-Bf   (repet.py:837, 852)
        slot_frames = np.where(slots <= j0, j - (j0 - slots), j - (j0 - slots) - buffer_length2)
        buffer_spectrogram = magnitude[:, slot_frames, :]
        similarity_vector = similaritymatrix(
            np.mean(buffer_spectrogram, axis=2), np.mean(buffer_spectrogram[:, j0 : j0 + 1, :], axis=2)
        )
        _, similarity_indices = localmaxima(
            similarity_vector[:, 0], cfg["similarity_threshold"], similarity_distance2, cfg["similarity_number"]
        )
        all_indices.append(slot_frames[similarity_indices])
        k = j * step_length
        for i in range(number_channels):
            with np.errstate(invalid="ignore"):
                repeating_spectrum = np.median(buffer_spectrogram[:, similarity_indices, i], axis=1)
            repeating_spectrum = np.minimum(repeating_spectrum, buffer_spectrogram[:, j0, i])
            repeating_mask = (repeating_spectrum + EPS) / (buffer_spectrogram[:, j0, i] + EPS)
            repeating_mask[1 : cutoff2 + 1] = 1
            repeating_mask = np.concatenate((repeating_mask, repeating_mask[-2:0:-1]))
            background_signal[k : k + window_length, i] += np.real(
                np.fft.ifft(repeating_mask * frame_ft[:, j, i], axis=0)
            )
    background_signal = background_signal[0:number_samples, :]
    background_signal = background_signal / sum(window_function[0:window_length:step_length])
    if return_details:
        return background_signal, dict(indices=all_indices, first_frame=buffer_length2 - 1)
    return background_signal
