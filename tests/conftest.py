"""Shared pytest configuration: markers, import paths, golden-vector loaders."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "repet-python_b200"), ROOT):
    if path not in sys.path:
        sys.path.insert(0, path)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_helpers():
    return dict(np.load(os.path.join(GOLDEN, "helpers.npz")))


@pytest.fixture(scope="session")
def golden_drivers():
    return dict(np.load(os.path.join(GOLDEN, "drivers.npz")))


@pytest.fixture(scope="session")
def golden_rates():
    """Reference outputs at 8, 16, 22.05 and 48 kHz (window lengths 512, 1024, 2048)."""
    return dict(np.load(os.path.join(GOLDEN, "drivers_rates.npz")))


@pytest.fixture(scope="session")
def wav_pcm():
    data = np.load(os.path.join(GOLDEN, "audio_file_int16.npz"))
    assert int(data["sampling_frequency"]) == 44100
    return data["pcm"]
