import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(ROOT, "guitar-tablature-classification_b200")
for p in (ROOT, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """libgtc.so, built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()                      # incremental: recompiles only sources newer than their objects
    from gtc_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def basis_cache():
    return {}


@pytest.fixture(scope="session")
def recipe():
    from gtc_b200 import CqtRecipe
    return CqtRecipe()


def make_test_audio(n_samples, seed, sr=22050.0):
    """Cheap deterministic 'guitar-ish' audio for parity tests: decaying harmonic tones + noise, float32."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / sr
    y = 0.003 * rng.standard_normal(n_samples)
    for _ in range(max(1, int(6 * n_samples / sr))):
        f = 440.0 * 2 ** ((rng.uniform(40, 82) - 69) / 12)
        on = rng.uniform(0, n_samples / sr)
        rel = np.maximum(t - on, 0)
        env = np.where(t >= on, np.exp(-rel / rng.uniform(0.15, 0.7)), 0.0)
        y += rng.uniform(0.3, 1.0) * env * (np.sin(2 * np.pi * f * rel) + 0.5 * np.sin(4 * np.pi * f * rel))
    y *= 0.5 / np.abs(y).max()
    return y.astype(np.float32)
