"""LMS notch / noise reduction: the oracle's restatement against the reference's own block compiled on the
host (oracle/_ref, one stream per fresh process: its state is in function statics) and against the golden vectors.  The float
path turns out bit-exact (every operation a separately rounded IEEE operation in source order), so no tolerance is needed."""
import os

import numpy as np
import pytest

import anr_lib as al
import oracle_lib as ol

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "anr_kat.npz")
needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("mode", [1, 2])
def test_anr_matches_compiled_reference(mode):
    x = al.audio_stream(3, 128 * 60, seed=mode)
    x[2, 500:900] = 32767  # a burst at full scale
    x[2, 900:1300] = -32768
    o = al.OrcAnr(3)
    y = o.run(mode, x)
    for c in range(3):
        assert np.array_equal(y[c], al.ref_anr_run(mode, x[c])), (mode, c)
    assert np.abs(y.astype(np.int32) - x).max() > 100  # the filter does something


def test_anr_golden():
    z = np.load(G)
    for mode in (1, 2):
        assert np.array_equal(al.OrcAnr(2).run(mode, z["x"]), z[f"y_mode{mode}"])


def test_anr_state_carry_and_notch_effect():
    x = al.audio_stream(1, 128 * 200, seed=9)
    o1, o2 = al.OrcAnr(1), al.OrcAnr(1)
    whole = o1.run(1, x)
    parts = np.concatenate([o2.run(1, x[:, i:i + 128 * 8]) for i in range(0, x.shape[1], 128 * 8)], axis=1)
    assert np.array_equal(whole, parts)
    # the steady tone is attenuated once the LMS has converged
    f0 = 0.031
    t = np.arange(x.shape[1])
    def tone_power(v):
        seg = v[0, -4096:].astype(np.float64)
        return abs(np.dot(seg, np.exp(-2j * np.pi * f0 * t[-4096:]))) / 4096
    assert tone_power(whole) < 0.25 * tone_power(x)
