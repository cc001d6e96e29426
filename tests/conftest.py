import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def orc():
    import oracle_lib
    return oracle_lib.CheckerLib("orc")


@pytest.fixture(scope="session")
def ref():
    import oracle_lib
    if not oracle_lib.have_ref():
        pytest.skip("oracle/_ref/libmsdr_ref.so not available (needs /root/reference at build time)")
    return oracle_lib.CheckerLib("ref")


@pytest.fixture(scope="session")
def K():
    import minimal_sdr_b200 as m
    return m.load_ref_constants()


@pytest.fixture(scope="session")
def msdr():
    """The product package; GPU tests call through its ctypes binding of the C ABI."""
    import minimal_sdr_b200 as m
    m.capi.lib()
    return m


def adversarial_inputs(n, rng):
    """Input streams that exercise wrap / saturation corners (SURVEY.md 8d 'adversarial set')."""
    k = np.arange(n)
    return {
        "uniform": rng.integers(-32768, 32768, n, dtype=np.int16),
        "all_min": np.full(n, -32768, np.int16),
        "all_max": np.full(n, 32767, np.int16),
        "alt_fs4": np.array([32767, 32767, -32767, -32767], np.int16)[k % 4],
        "alt_fs4_min": np.array([32767, -32768, -32768, 32767], np.int16)[k % 4],
        "square_fs8": np.where((k // 4) % 2 == 0, 32767, -32768).astype(np.int16),
        "impulse": np.where(k == 5, 32767, 0).astype(np.int16),
        "zeros": np.zeros(n, np.int16),
    }


def wrap_coeffs(T, rng):
    """Taps with sum|c| >> 65536: forces the 32-bit accumulator to wrap and the output to saturate."""
    return rng.integers(-32768, 32768, T, dtype=np.int16)
