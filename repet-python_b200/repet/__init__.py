"""
repet -- drop-in, B200-native replacement for zafarrafii/REPET-Python's `repet.py`.

Same module surface as the reference (repet.py:15-25, 42-63): the five separation
functions take and return NumPy arrays, the nine tunables are module globals read at call
time, `wavread` / `wavwrite` / `specshow` keep their signatures.  The arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of `include/repet_b200.h`
(librepet_b200.so, loaded by `_host.py` through ctypes).  There is no CPU fallback: without
the library or without a CUDA device every separation call raises.

Functions:
    original - Compute the original REPET.
    extended - Compute REPET extended.
    adaptive - Compute the adaptive REPET.
    sim - Compute REPET-SIM.
    simonline - Compute the online REPET-SIM.

Other:
    wavread - Read a WAVE file (using SciPy).
    wavwrite - Write a WAVE file (using SciPy).
    specshow - Display an spectrogram in dB, seconds, and Hz.

Beyond the reference: `*_batch` for many clips per call, `SimOnline` for block-wise
streaming, `separate` (background + foreground + the three display spectrograms of README.md:64-81 in one
device pass), `spectrogram` and `spectrogram_db`.
"""

import numpy as np

from . import _host

# Public variables (repet.py:42-63) -- read at call time, so `repet.period_range = [1, 5]` works
# Cutoff frequency in Hz for the dual high-pass filter of the foreground
cutoff_frequency = 100

# Period range in seconds for the beat spectrum (original, extended, adaptive)
period_range = [1, 10]

# Segment length and step in seconds (extended, adaptive)
segment_length = 10
segment_step = 5

# Filter order for the median filter (adaptive)
filter_order = 5

# Minimal threshold for two similar frames in [0,1], minimal distance between two similar frames in
# seconds, and maximal number of similar frames for every frame (sim, simonline)
similarity_threshold = 0
similarity_distance = 1
similarity_number = 100

# Buffer length in seconds (simonline)
buffer_length = 10

_TUNABLES = (
    "cutoff_frequency",
    "period_range",
    "segment_length",
    "segment_step",
    "filter_order",
    "similarity_threshold",
    "similarity_distance",
    "similarity_number",
    "buffer_length",
)


def _tunables():
    module = globals()
    return {name: module[name] for name in _TUNABLES}


# ----------------------------------------------------------------------------------------
# Public functions
# ----------------------------------------------------------------------------------------
def original(audio_signal, sampling_frequency):
    """
    Compute the original REPET (repet.py:67-202).

    Inputs:
        audio_signal: audio signal (number_samples, number_channels)
        sampling_frequency: sampling frequency in Hz
    Output:
        background_signal: background signal (number_samples, number_channels)
    """
    return _host.original_f64(audio_signal, sampling_frequency, _tunables())


def original_batch(audio_signals, sampling_frequency, devices=None):
    """
    The original REPET over a batch of equally long clips in one call.

    Inputs:
        audio_signals: float32 array (number_clips, number_channels, number_samples), planar
        sampling_frequency: sampling frequency in Hz
    Outputs:
        background_signals: float32 array of the same shape
        repeating_periods: int32 array (number_clips,), in time frames

    A list / tuple of (number_channels, number_samples_i) arrays of DIFFERENT lengths is accepted too: the clips
    are grouped by shape (never padded: a clip's period range depends on its length) and the outputs come back as
    a list of backgrounds and an int32 array of periods, in input order.

    `devices` (e.g. range(8)) shards the batch by clip over several GPUs inside this process: one handle and one
    host thread per GPU, no collective (clips are independent).
    """
    if isinstance(audio_signals, (list, tuple)):
        backgrounds, periods = _host.driver_batch_ragged("original", audio_signals, sampling_frequency, _tunables())
        return backgrounds, np.array([int(p[0]) for p in periods], dtype=np.int32)
    if devices is not None:
        background, periods = _host.separate_batch("original", audio_signals, sampling_frequency, _tunables(), devices=devices)
        return background, periods[:, 0]
    return _host.original_batch(audio_signals, sampling_frequency, _tunables())


def separate_batch(audio_signals, sampling_frequency, method="original", in_format="f32", out_format="f32", out=None,
                   devices=None):
    """
    Any of the five methods over a batch of equally long clips, with the sample format of either side chosen by
    the caller (the steps either side of the separation in the reference's usage, repet.py:914-946):

        in_format  "f32": (number_clips, number_channels, number_samples) float32, planar
                   "pcm16": (number_clips, number_samples, number_channels) int16 as scipy.io.wavfile.read returns
                            it; normalised by 2^15 on the device, as repet.wavread does (repet.py:929)
        out_format "f32" or "pcm16" (round(y * 2^15) saturated to int16: lossy, for int16 WAVE writers)
        devices    optional GPU indices: clip-sharded over them inside this process (one thread per GPU)

    Outputs: background_signals in `out_format`; integers int32 (number_clips, per clip: 1 period | segment periods |
    frame periods | [counts][indices] lists).  `repet.pinned_empty` allocates page-locked arrays for `audio_signals`
    and `out`, which the copies need to run at the full rate of the host link.
    """
    return _host.separate_batch(method, audio_signals, sampling_frequency, _tunables(), in_format=in_format,
                                out_format=out_format, out=out, devices=devices)


def separate_files(input_files, output_files, method="original", devices=None):
    """
    WAVE files in, WAVE files out: the steps either side of the separation in the reference's usage (wavread ->
    method -> wavwrite, README.md:56-68) with the samples never leaving their 16-bit form on the host.

    int16 files go to the device as int16 PCM (normalised by 2^15 there, as repet.wavread does, repet.py:929) and come
    back as int16 (round(y * 2^15), saturated); files of equal length, channel count and sampling rate share one
    batch call (`devices` shards each batch over several GPUs).  Files in another sample format are read by
    `wavread`, separated through the float64 entry point and written back as float64 by `wavwrite`, exactly as the
    reference's own example does.  Returns the list of integer outputs (periods or lists) per file.
    """
    import scipy.io.wavfile

    input_files, output_files = list(input_files), list(output_files)
    if len(input_files) != len(output_files):
        raise ValueError("one output file per input file expected")
    loaded = [scipy.io.wavfile.read(path) for path in input_files]
    integers = [None] * len(loaded)
    groups = {}
    for index, (sampling_frequency, data) in enumerate(loaded):
        if data.dtype == np.int16:
            pcm = data if data.ndim == 2 else data[:, np.newaxis]
            groups.setdefault((int(sampling_frequency), pcm.shape), []).append((index, pcm))
        else:  # the reference's own route
            audio_signal = data / pow(2, data.itemsize * 8 - 1) if data.dtype.kind in "iu" else data.astype(float)
            if audio_signal.ndim == 1:
                audio_signal = audio_signal[:, np.newaxis]
            result = _host.separate_f64(method, audio_signal, sampling_frequency, _tunables(), spectrograms=False)
            wavwrite(result["background"], sampling_frequency, output_files[index])
            integers[index] = result["integers"]
    for (sampling_frequency, _), members in groups.items():
        batch = np.stack([pcm for _, pcm in members])
        background, ints = _host.separate_batch(method, batch, sampling_frequency, _tunables(), in_format="pcm16",
                                                out_format="pcm16", devices=devices)
        for slot, (index, _) in enumerate(members):
            scipy.io.wavfile.write(output_files[index], sampling_frequency, background[slot])
            integers[index] = ints[slot]
    return integers


def pinned_empty(shape, dtype=np.float32):
    """An uninitialised NumPy array on page-locked host memory; keep the returned holder alive while the array is
    in use: `holder = repet.pinned_empty(shape); x = holder.array`."""
    return _host.PinnedArray(shape, dtype)


def original_batch_pcm16(pcm_signals, sampling_frequency):
    """The original REPET over int16 PCM clips as stored in WAVE files: (number_clips, number_samples,
    number_channels) int16 -> float32 backgrounds (number_clips, number_channels, number_samples), periods."""
    return _host.original_batch_pcm16(pcm_signals, sampling_frequency, _tunables())


def extended_batch(audio_signals, sampling_frequency):
    """REPET extended over a batch (see original_batch); periods: int32 (number_clips, number_segments)."""
    return _host.driver_batch("extended", audio_signals, sampling_frequency, _tunables())


def adaptive_batch(audio_signals, sampling_frequency):
    """The adaptive REPET over a batch (see original_batch); periods: int32 (number_clips, number_times)."""
    return _host.driver_batch("adaptive", audio_signals, sampling_frequency, _tunables())


def sim_batch(audio_signals, sampling_frequency):
    """REPET-SIM over a batch (see original_batch); lists: int32 (number_clips, number_times*(similarity_number+1)),
    per clip [counts][indices] -- repet._host.unpack_lists turns a row into the reference's ragged list."""
    return _host.driver_batch("sim", audio_signals, sampling_frequency, _tunables())


def extended(audio_signal, sampling_frequency, devices=None):
    """Compute REPET extended (repet.py:205-419).  `devices` (beyond the reference): split ONE long track by time
    block over several GPUs (whole segments per block, junctions cross-faded as the reference does)."""
    if devices is not None:
        return _host.sharded_track("extended", audio_signal, sampling_frequency, _tunables(), devices)
    return _host.extended_f64(audio_signal, sampling_frequency, _tunables())


def adaptive(audio_signal, sampling_frequency, devices=None):
    """Compute the adaptive REPET (repet.py:422-568).  `devices`: ONE long track by time block over several GPUs
    (blocks cut on the beat-spectrogram grid, separated with a halo)."""
    if devices is not None:
        return _host.sharded_track("adaptive", audio_signal, sampling_frequency, _tunables(), devices)
    return _host.adaptive_f64(audio_signal, sampling_frequency, _tunables())


def sim(audio_signal, sampling_frequency):
    """Compute REPET-SIM (repet.py:571-709)."""
    return _host.sim_f64(audio_signal, sampling_frequency, _tunables())


def simonline(audio_signal, sampling_frequency, devices=None):
    """Compute the online REPET-SIM (repet.py:712-911).  `devices`: ONE long track by time block over several GPUs
    (every block replays the similarity history of its first frames)."""
    if devices is not None:
        return _host.sharded_track("simonline", audio_signal, sampling_frequency, _tunables(), devices)
    return _host.simonline_f64(audio_signal, sampling_frequency, _tunables())


def separate(audio_signal, sampling_frequency, method="original", spectrograms=True):
    """
    The documented usage of the reference in one call (README.md:64-81): estimate the background with
    `method` ("original", "extended", "adaptive", "sim", "simonline"), the foreground as audio - background,
    and the mixture / background / foreground spectrograms abs(_stft(mean(x, axis=1)))[0:number_frequencies],
    all on the device from buffers already resident.

    Output: dict with background, foreground (number_samples, number_channels); audio_spectrogram,
    background_spectrogram, foreground_spectrogram (number_frequencies, number_times); integers.
    """
    return _host.separate_f64(method, audio_signal, sampling_frequency, _tunables(), spectrograms=spectrograms)


def spectrogram(audio_signal, sampling_frequency):
    """Magnitude spectrogram of the channel mean as the reference's examples compute it (README.md:79):
    abs(_stft(mean(audio_signal, axis=1), hamming, N/2))[0:N/2+1] -> (number_frequencies, number_times)."""
    audio_signal = np.asarray(audio_signal, dtype=float)
    if audio_signal.ndim == 1:
        audio_signal = audio_signal[:, np.newaxis]
    params, window_function = _host.derive_params(sampling_frequency, _tunables())
    half = _host.stft_half(np.mean(audio_signal, axis=1)[np.newaxis, :], window_function, params.step_length)[0]
    return np.abs(half).T


def spectrogram_db(audio_spectrogram):
    """The image `specshow` displays (repet.py:982): 20*log10 of a magnitude spectrogram."""
    return 20 * np.log10(audio_spectrogram)


class SimOnline:
    """Streaming front end of the online REPET-SIM: `SimOnline(sampling_frequency, number_channels)`, then
    `process(block)` per block of samples and `flush()` at the end; the concatenated outputs equal
    `repet.simonline` on the whole signal.  The module tunables are read at construction.

    Streams the fast kernels cover (window lengths up to 2048, one or two channels) run on the stateful C stream
    (sample history resident on the device); other sampling rates and channel counts use the same bookkeeping on the
    host over the general float64 device path."""

    def __init__(self, sampling_frequency, number_channels):
        tunables = _tunables()
        params, _ = _host.derive_params(sampling_frequency, tunables, "simonline")
        if _host.needs_general_path(params, int(number_channels)):
            self._impl = _host.SimOnlineStreamHost(sampling_frequency, number_channels, tunables)
        else:
            self._impl = _host.SimOnlineStream(sampling_frequency, number_channels, tunables)

    def process(self, block):
        return self._impl.process(block)

    def flush(self):
        return self._impl.flush()

    def close(self):
        if hasattr(self._impl, "close"):
            self._impl.close()


def wavread(audio_file):
    """
    Read a WAVE file (using SciPy) (repet.py:914-931).

    Input:
        audio_file: path to an audio file
    Outputs:
        audio_signal: audio signal (number_samples, number_channels)
        sampling_frequency: sampling frequency in Hz
    """
    import scipy.io.wavfile

    sampling_frequency, audio_signal = scipy.io.wavfile.read(audio_file)
    audio_signal = audio_signal / pow(2, audio_signal.itemsize * 8 - 1)
    return audio_signal, sampling_frequency


def wavwrite(audio_signal, sampling_frequency, audio_file):
    """Write a WAVE file (using SciPy) (repet.py:934-946)."""
    import scipy.io.wavfile

    scipy.io.wavfile.write(audio_file, sampling_frequency, audio_signal)


def specshow(audio_spectrogram, time_duration, maximum_frequency, xtick_step=1, ytick_step=1000):
    """
    Display a spectrogram in dB, seconds, and Hz (repet.py:949-997).  matplotlib is imported
    lazily: it is a plotting convenience, not part of the separation path.
    """
    import matplotlib.pyplot as plt

    number_frequencies, number_times = np.shape(audio_spectrogram)
    time_resolution = number_times / time_duration
    frequency_resolution = number_frequencies / maximum_frequency
    xtick_locations = np.arange(xtick_step * time_resolution, number_times, xtick_step * time_resolution)
    xtick_labels = np.arange(xtick_step, time_duration, xtick_step).astype(int)
    ytick_locations = np.arange(ytick_step * frequency_resolution, number_frequencies, ytick_step * frequency_resolution)
    ytick_labels = np.arange(ytick_step, maximum_frequency, ytick_step).astype(int)
    plt.imshow(spectrogram_db(audio_spectrogram), aspect="auto", cmap="jet", origin="lower")
    plt.xticks(ticks=xtick_locations, labels=xtick_labels)
    plt.yticks(ticks=ytick_locations, labels=ytick_labels)
    plt.xlabel("Time (s)")
    plt.ylabel("Frequency (Hz)")


# ----------------------------------------------------------------------------------------
# Private functions with the reference's shapes (repet.py:1001-1545)
# ----------------------------------------------------------------------------------------
def _stft(audio_signal, window_function, step_length):
    """
    Short-time Fourier transform (repet.py:1001-1060).

    Inputs:
        audio_signal: audio signal (number_samples,)
        window_function: window function (window_length,)
        step_length: step length in samples
    Output:
        audio_stft: audio STFT (window_length, number_times), full mirrored spectrum
    """
    audio_signal = np.asarray(audio_signal)
    if not _host.fast_stft_applies(len(window_function), step_length):
        # any window length and step (repet.py:1018-1058 puts no constraint on them): float64 general path
        return _host.stft_general(audio_signal, window_function, step_length)
    half = _host.stft_half(audio_signal[np.newaxis, :], window_function, step_length)[0]  # (T, F)
    window_length = len(window_function)
    audio_stft = np.empty((window_length, half.shape[0]), dtype=complex)
    audio_stft[0 : window_length // 2 + 1, :] = half.T
    audio_stft[window_length // 2 + 1 :, :] = np.conj(half[:, window_length // 2 - 1 : 0 : -1].T)
    return audio_stft


def _istft(audio_stft, window_function, step_length):
    """
    Inverse short-time Fourier transform (repet.py:1063-1105): real(ifft) of every frame,
    overlap-add, trim, divide by the COLA gain.

    Inputs:
        audio_stft: audio STFT (window_length, number_times)
        window_function: window function (window_length,)
        step_length: step length in samples
    Output:
        audio_signal: audio signal (number_samples,)
    """
    audio_stft = np.asarray(audio_stft, dtype=complex)
    window_length = audio_stft.shape[0]
    if not _host.fast_stft_applies(window_length, step_length) or len(window_function) != window_length:
        return _host.istft_general(audio_stft, window_function, step_length)
    half = window_length // 2
    # real(ifft(Y)) only sees the Hermitian part of Y: Yh[k] = (Y[k] + conj(Y[N-k]))/2
    mirror = np.conj(audio_stft[(-np.arange(window_length)) % window_length, :])
    hermitian = 0.5 * (audio_stft + mirror)
    signal = _host.istft_half(hermitian[np.newaxis, 0 : half + 1, :].transpose(0, 2, 1), window_function, step_length)
    return signal[0].astype(float)


def _acorr(data_matrix):
    """Autocorrelation of every column using the Wiener-Khinchin theorem (repet.py:1108-1139)."""
    return _host.acorr(data_matrix)


def _beatspectrum(audio_spectrogram):
    """Beat spectrum (repet.py:1142-1158): (number_frequencies, number_times) -> (number_times,)."""
    return _host.beatspectrum(audio_spectrogram)


def _beatspectrogram(audio_spectrogram, segment_length, segment_step):
    """Beat spectrogram (repet.py:1161-1206): (number_frequencies, number_times) -> (segment_length, number_times)."""
    return _host.beatspectrogram(audio_spectrogram, segment_length, segment_step)


def _selfsimilaritymatrix(data_matrix):
    """Self-similarity matrix using the cosine similarity (repet.py:1209-1225), exact float64."""
    return _host.similaritymatrix(data_matrix, data_matrix)


def _similaritymatrix(data_matrix1, data_matrix2):
    """Similarity matrix using the cosine similarity (repet.py:1228-1246), exact float64."""
    return _host.similaritymatrix(data_matrix1, data_matrix2)


def _localmaxima(data_vector, minimum_value, minimum_distance, number_values):
    """Values and indices of the local maxima in a vector (repet.py:1294-1345)."""
    return _host.localmaxima(data_vector, minimum_value, minimum_distance, number_values)


def _indices(similarity_matrix, similarity_threshold, similarity_distance, similarity_number):
    """Similarity indices from the similarity matrix (repet.py:1348-1383)."""
    return _host.indices(similarity_matrix, similarity_threshold, similarity_distance, similarity_number)


def _simmask(audio_spectrogram, similarity_indices):
    """Repeating mask for REPET-SIM (repet.py:1511-1545)."""
    return _host.simmask(audio_spectrogram, similarity_indices)


def _periods(beat_spectrogram, period_range):
    """Repeating period(s) from a beat spectrum or beat spectrogram (repet.py:1249-1291)."""
    return _host.periods(beat_spectrogram, period_range)


def _adaptivemask(audio_spectrogram, repeating_periods, filter_order):
    """Repeating mask for the adaptive REPET (repet.py:1461-1508)."""
    return _host.adaptivemask(audio_spectrogram, repeating_periods, filter_order)


def _mask(audio_spectrogram, repeating_period):
    """Repeating mask for REPET (repet.py:1386-1458): (number_frequencies, number_times), period ->
    (number_frequencies, number_times)."""
    return _host.mask(audio_spectrogram, repeating_period)
