"""Synchronous-AM PLL demodulator: the oracle's restatement against the reference's own arm compiled on the
host (both use this host's libm, so they agree bit for bit) and the golden vectors; the library's loop constants."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import syncam_lib as sl

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "syncam_kat.npz")
needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
def test_syncam_matches_compiled_reference():
    I, Q = sl.baseband(3, 128 * 120, seed=2)
    I[2, 4000:4600] = 32767  # a burst at full scale
    y = sl.OrcSyncAm(3).run(I, Q)
    for c in range(3):
        assert np.array_equal(y[c], sl.ref_syncam_run(I[c], Q[c])), c
    assert np.abs(y.astype(np.int32)).max() > 1000


def test_syncam_golden():
    """Vectors made from the compiled reference on the build host.  libm may differ in the last ulp on another host: tolerance as
    for the GPU (relative RMS <= 1e-5, at most 1 LSB anywhere)."""
    z = np.load(G)
    y = sl.OrcSyncAm(z["I"].shape[0]).run(z["I"], z["Q"]).astype(np.float64)
    ref = z["y"].astype(np.float64)
    assert np.abs(y - ref).max() <= 1
    assert np.sqrt(np.mean((y - ref) ** 2)) <= 1e-5 * np.sqrt(np.mean(ref ** 2))


def test_library_constants_match_the_oracle(msdr):
    """msdr_syncam_constants evaluates the sketch's static initialisers in C++ (float exp overload): same bits as the oracle's."""
    a = np.array(msdr.syncam.constants(), np.float32)
    b = np.array(sl.OrcSyncAm(1).constants(), np.float32)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_pll_locks():
    """After lock the output follows the envelope: corr[0] ~ |I + jQ| (the carrier offset is removed by the PLL)."""
    I, Q = sl.baseband(1, 128 * 300, seed=4)
    y = sl.OrcSyncAm(1).run(I, Q)[0].astype(np.float64)
    env = np.hypot(I[0].astype(np.float64), Q[0].astype(np.float64))
    tail = slice(-8192, None)
    assert np.corrcoef(y[tail], env[tail])[0, 1] > 0.99
