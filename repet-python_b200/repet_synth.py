"""
Seeded synthetic audio for the REPET benchmarks and parity tests.

A clip is a periodically repeating background (12 Hann-enveloped bursts per period, every
third one a broadband noise hit, the others sinusoids with an independent phase per
channel), a NON-periodic foreground (a tone on a random-walk frequency, gated on and off)
and a noise floor, so that no frame is digital silence (reference quirk Q18) and there
are no exact ties.  The master copy is float32; the CPU oracle reads a float64 view of
the same samples.

This is measurement/test input only: nothing here is on the separation path.
"""

import numpy as np


def _background_period(rng, period_samples, number_channels, sampling_frequency):
    period = np.zeros((number_channels, period_samples), dtype=np.float64)
    for burst in range(12):
        length = int(rng.uniform(0.05, 0.25) * sampling_frequency)
        length = max(8, min(length, period_samples))
        amplitude = rng.uniform(0.05, 0.2)
        start = int(rng.integers(0, period_samples - length + 1))
        envelope = np.hanning(length)
        if burst % 3 == 2:
            body = rng.standard_normal((number_channels, length))
        else:
            frequency = rng.uniform(80.0, 4000.0)
            phase = rng.uniform(0.0, 2.0 * np.pi, size=(number_channels, 1))
            body = np.sin(2.0 * np.pi * frequency * np.arange(length)[None, :] / sampling_frequency + phase)
        period[:, start : start + length] += amplitude * envelope[None, :] * body
    return period


def make_clip(
    index,
    number_samples,
    number_channels=2,
    sampling_frequency=44100,
    step_length=1024,
    redraw_seconds=None,
    dtype=np.float32,
):
    """Return clip `index` as an array of shape (number_channels, number_samples) (planar).

    `redraw_seconds=(lo, hi)` re-draws the background period every U[lo, hi] seconds (used
    for the long tracks of BASELINE configs 3-5, so that adaptive/extended have something
    to adapt to)."""
    rng = np.random.default_rng(1000 + int(index))
    fs = float(sampling_frequency)
    out = np.zeros((number_channels, number_samples), dtype=np.float64)

    # repeating background
    position = 0
    while position < number_samples:
        period_samples = int(rng.uniform(1.2, 4.0) * fs)
        if period_samples % step_length == 0:
            period_samples += 1
        if redraw_seconds is None:
            span = number_samples - position
        else:
            span = min(number_samples - position, int(rng.uniform(*redraw_seconds) * fs))
        period = _background_period(rng, period_samples, number_channels, fs)
        repeats = -(-span // period_samples)
        out[:, position : position + span] += np.tile(period, (1, repeats))[:, :span]
        position += span

    # non-periodic foreground: random-walk tone, gated
    control_rate = 100.0
    number_control = int(np.ceil(number_samples / fs * control_rate)) + 2
    walk = np.cumsum(rng.standard_normal(number_control))
    walk = (walk - walk.min()) / max(walk.max() - walk.min(), 1e-12)
    frequency = 200.0 + 600.0 * walk
    time_control = np.arange(number_control) / control_rate
    time_samples = np.arange(number_samples) / fs
    instantaneous = np.interp(time_samples, time_control, frequency)
    phase = 2.0 * np.pi * np.cumsum(instantaneous) / fs
    gate_rate = 2.0
    number_gate = int(np.ceil(number_samples / fs * gate_rate)) + 2
    gate = (rng.uniform(size=number_gate) < 0.6).astype(np.float64)
    gate_samples = np.interp(time_samples, np.arange(number_gate) / gate_rate, gate)
    tone = 0.08 * np.sin(phase) * gate_samples
    gains = np.array([1.0, 0.8] + [0.9] * max(0, number_channels - 2))[:number_channels]
    out += gains[:, None] * tone[None, :]

    # noise floor
    out += 0.005 * rng.standard_normal((number_channels, number_samples))
    return out.astype(dtype)


def _clip_job(args):
    index, number_samples, number_channels, sampling_frequency, kwargs = args
    return make_clip(index, number_samples, number_channels, sampling_frequency, **kwargs)


def make_batch(first_index, number_clips, number_samples, number_channels=2, sampling_frequency=44100,
               out=None, workers=None, processes=False, **kwargs):
    """Clips first_index .. first_index+number_clips-1 as (B, C, S) float32.

    `processes=True` uses a fork pool (call it BEFORE CUDA is initialised in the process);
    the default is a thread pool."""
    import os

    if out is None:
        out = np.empty((number_clips, number_channels, number_samples), dtype=np.float32)
    workers = max(1, workers or min(64, os.cpu_count() or 1))
    jobs = [(first_index + i, number_samples, number_channels, sampling_frequency, kwargs) for i in range(number_clips)]
    if processes and workers > 1 and number_clips > 1:
        import multiprocessing

        with multiprocessing.get_context("fork").Pool(min(workers, number_clips)) as pool:
            for i, clip in enumerate(pool.imap(_clip_job, jobs, chunksize=1)):
                out[i] = clip
    else:
        from concurrent.futures import ThreadPoolExecutor

        def fill(i):
            out[i] = _clip_job(jobs[i])

        with ThreadPoolExecutor(max_workers=workers) as pool:
            list(pool.map(fill, range(number_clips)))
    return out
