"""Front-end conditioning: the oracle's restatement against the reference's own code compiled on the host
(oracle/_ref) and against the committed golden vectors made from it.  CPU only."""
import os

import numpy as np
import pytest

import frontend_lib as fl
import oracle_lib as ol

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend_kat.npz")
needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


@pytest.fixture(scope="module")
def orc():
    return fl.Orc()


@pytest.fixture(scope="module")
def ref():
    return fl.Ref()


@needs_ref
def test_hpf_matches_compiled_reference(orc, ref):
    rng = np.random.default_rng(1)
    streams = [fl.adc_stream(1, 128 * 40, seed=2)[0],
               rng.integers(0, 65536, 128 * 20).astype(np.uint16),               # full 16-bit codes: the accumulator wraps
               np.r_[np.zeros(128 * 3), np.full(128 * 3, 65535), np.zeros(128 * 2)].astype(np.uint16),  # rails: output saturates
               np.full(128 * 4, 2048, np.uint16)]
    for s in streams:
        for x1, y1 in ((0, 0), (int(s[0]) << 14, 0), (123456789, -987654321)):
            a = orc.hpf(s, x1, y1)
            b = ref.hpf(s, x1, y1)
            assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
    # state carry: block by block == one call
    s = streams[0]
    whole = orc.hpf(s)
    x1 = y1 = 0
    parts = []
    for i in range(0, s.size, 128):
        o, x1, y1 = orc.hpf(s[i:i + 128], x1, y1)
        parts.append(o)
    assert np.array_equal(np.concatenate(parts), whole[0]) and (x1, y1) == whole[1:]


@needs_ref
def test_amplifier_matches_compiled_reference(orc, ref):
    rng = np.random.default_rng(3)
    blk = rng.integers(-32768, 32768, 128).astype(np.int16)
    blk[:4] = [-32768, 32767, 0, -1]
    for gain in (0.25, 1.0, 0.999999, 1.5, 40.0, 3.3e-6, 0.0, -2.0, 1e9, -1e9, 0.1, 17.123):
        assert orc.amp_multiplier(gain) == ref.amp_multiplier(gain)
        a, sa = orc.amp_apply(blk, orc.amp_multiplier(gain))
        b, sb = ref.amp_block(gain, blk)
        assert sa == sb
        if sb:  # the reference transmits nothing at multiplier 0; the oracle then writes zeros
            assert np.array_equal(a, b), gain
        else:
            assert not a.any()


def _agc_blocks(seed, n):
    rng = np.random.default_rng(seed)
    levels = np.abs(rng.normal(0, 1, n)) * rng.choice([30, 300, 3000, 12000, 16000, 20000, 30000, 40000], n)
    blocks = (rng.normal(0, 1, (n, 128)) * levels[:, None] * 0.4).clip(-32768, 32767).astype(np.int16)
    blocks[5] = 0
    blocks[6, 7] = -32768
    blocks[7] = 32767
    blocks[8, ::2] = -32768
    return blocks


@needs_ref
@pytest.mark.parametrize("seed,start,mx", [(0, 0.25, 40.0), (1, 5.0, 40.0), (2, 39.0, 40.0), (3, 0.05, 1.0)])
def test_agc_trajectory_matches_compiled_reference(orc, seed, start, mx):
    """Gain law, 25-block history incl. the dropped 26th store, absmax idiom: every block's AGC_val and multiplier."""
    blocks = _agc_blocks(seed, 400)
    vo, mo = orc.agc_trajectory(blocks, start, mx)
    vr, mr = fl.Ref.agc_trajectory(blocks, start, mx)
    assert np.array_equal(vo.view(np.uint32), vr.view(np.uint32))
    assert np.array_equal(mo, mr)
    assert len(set(mo.tolist())) > 5  # the gain really moves


def test_frontend_golden(orc):
    """Golden vectors made from the compiled reference (tests/golden/make_golden.py)."""
    z = np.load(G)
    out, x1, y1 = orc.hpf(z["hpf_in"], int(z["hpf_state"][0]), int(z["hpf_state"][1]))
    assert np.array_equal(out, z["hpf_out"]) and [x1, y1] == [int(v) for v in z["hpf_state_out"]]
    for g, m, o in zip(z["amp_gains"], z["amp_mults"], z["amp_out"]):
        assert orc.amp_multiplier(g) == int(m)
        assert np.array_equal(orc.amp_apply(z["amp_in"], int(m))[0], o)
    v, m = orc.agc_trajectory(z["agc_blocks"], 0.25, 40.0)
    assert np.array_equal(v.view(np.uint32), z["agc_val"].view(np.uint32)) and np.array_equal(m, z["agc_mult"])


def test_frontend_run_is_the_composition(orc):
    """orc_frontend_run == HPF -> amplifier -> AGC block by block, with per-channel state."""
    codes = fl.adc_stream(3, 128 * 60, seed=5)
    f = orc.frontend(3)
    f.preset(1, codes[1, 0])
    got = f.run(codes)
    for c in range(3):
        x1, y1 = ((int(codes[1, 0]) << 14, 0) if c == 1 else (0, 0))
        mult = orc.amp_multiplier(0.25)
        a = fl.OrcAgc()
        orc.L.orc_agc_init(a, 0.25, 40.0, 1)
        import ctypes as C
        for b in range(60):
            o, x1, y1 = orc.hpf(codes[c, b * 128:(b + 1) * 128], x1, y1)
            o, _ = orc.amp_apply(o, mult)
            assert np.array_equal(o, got[c, b * 128:(b + 1) * 128]), (c, b)
            m = C.c_int32(mult)
            if orc.L.orc_agc_update(C.byref(a), orc.L.orc_agc_absmax(np.ascontiguousarray(o)), C.byref(m)):
                mult = m.value
        st = f.state(c)
        assert (st["hpf_x1"], st["hpf_y1"], st["multiplier"]) == (x1, y1, mult)
