import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class _Golden:
    """npz fixtures with 'case/key' names -> nested access."""

    def __init__(self, path):
        self._z = np.load(path, allow_pickle=False)

    def case(self, name):
        pre = name + "/"
        return {k[len(pre):]: self._z[k] for k in self._z.files if k.startswith(pre)}

    def cases(self):
        return sorted({k.split("/")[0] for k in self._z.files})


@pytest.fixture(scope="session")
def golden_mfcc():
    return _Golden(os.path.join(GOLDEN, "mfcc_ref.npz"))


@pytest.fixture(scope="session")
def golden_gmm():
    return _Golden(os.path.join(GOLDEN, "gmm_ref.npz"))


@pytest.fixture(scope="session")
def golden_real():
    return _Golden(os.path.join(GOLDEN, "realparams_ref.npz"))


@pytest.fixture(scope="session")
def golden_silence():
    return _Golden(os.path.join(GOLDEN, "silence_ref.npz"))


def shipped_ubm_pools(golden_gmm, golden_real):
    """(w, mean, sigma) of the four UBMs the reference ships (src/gui/model/*.model), from the committed fixtures"""
    pools = [tuple(golden_gmm.case(t)[k] for k in ("w", "mean", "sigma")) for t in ("ubm32", "ubm64", "ubm256")]
    c = golden_real.case("ubm32p")
    return pools + [(c["w"], c["mean"], c["sigma"])]


def rel_err(a, ref):
    """the tolerance metric of the build: |a - ref| / max(1, |ref|)  (north_star: 1e-4 relative fp32;
    per-frame log-likelihoods can be ~0, SURVEY.md s7 hard part 2b)."""
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.abs(a - ref) / np.maximum(1.0, np.abs(ref))
