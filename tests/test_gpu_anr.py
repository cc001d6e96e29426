"""GPU parity of the LMS notch / noise reduction kernel (csrc/msdr_anr.cu, through the C ABI) against the CPU oracle, whose
restatement is pinned bit for bit to the reference's compiled block by tests/test_anr_cpu.py.  The kernel performs the same
separately rounded IEEE operations in the same order, so outputs AND the float state are required to be identical."""
import numpy as np
import pytest

import anr_lib as al
from chain_helpers import assert_same

pytestmark = pytest.mark.gpu


def _state_equal(gs, os_):
    assert np.array_equal(np.array(gs.w[:], np.float32).view(np.uint32), os_["w"].view(np.uint32))
    assert np.array_equal(np.array(gs.d[:], np.float32).view(np.uint32), os_["d"].view(np.uint32))
    assert np.float32(gs.lidx).view(np.uint32) == os_["lidx"].view(np.uint32)
    assert np.float32(gs.ngamma).view(np.uint32) == os_["ngamma"].view(np.uint32)
    assert gs.in_idx == os_["in_idx"]


@pytest.mark.parametrize("mode", [1, 2])
def test_anr_matches_oracle(msdr, mode):
    """37 channels (partial group), 70 blocks in ragged updates, state carried; outputs and float state bit-identical."""
    C, nb = 37, 70
    x = al.audio_stream(C, 128 * nb, seed=40 + mode)
    x[3, 1000:1500] = 32767
    x[4] = 0
    g = msdr.Anr(C)
    o = al.OrcAnr(C)
    outs, b0 = [], 0
    for n in (1, 5, 30, 2, 32):
        outs.append(g.update(mode, x[:, b0 * 128:(b0 + n) * 128]))
        b0 += n
    assert_same(np.concatenate(outs, axis=1), o.run(mode, x), f"anr mode {mode}")
    for c in (0, 3, 4, 36):
        _state_equal(g.get_state(c), o.state(c))


def test_anr_golden_and_state_migration(msdr):
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "anr_kat.npz"))
    for mode in (1, 2):
        assert_same(msdr.Anr(2).update(mode, z["x"]), z[f"y_mode{mode}"], f"anr golden mode {mode}")
    x = al.audio_stream(2, 128 * 24, seed=5)
    a = msdr.Anr(2)
    y1 = a.update(1, x[:, :128 * 12])
    b = msdr.Anr(2)
    for c in range(2):
        b.set_state(c, a.get_state(c))
    assert_same(np.concatenate([y1, b.update(1, x[:, 128 * 12:])], axis=1), al.OrcAnr(2).run(1, x), "anr migrated state")
    with pytest.raises(msdr.MsdrError):
        a.update(0, x)
