"""GPU parity of the front-end conditioning kernel (csrc/msdr_frontend.cu, through the C ABI) against the CPU oracle, whose
restatement is pinned to the reference's compiled code by tests/test_frontend_cpu.py.  Fixed point + IEEE float => bit-exact."""
import numpy as np
import pytest

import frontend_lib as fl
import oracle_lib as ol
from chain_helpers import assert_same, configure_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def forc():
    return fl.Orc()


def _state_equal(gs, os_):
    assert (gs.hpf_x1, gs.hpf_y1, gs.multiplier, gs.agc_idx) == (os_["hpf_x1"], os_["hpf_y1"], os_["multiplier"], os_["agc_idx"])
    assert np.float32(gs.agc_val).view(np.uint32) == os_["agc_val"].view(np.uint32)
    assert np.array_equal(np.array(gs.agc_buffer[:], np.int16), os_["agc_buffer"])


@pytest.mark.parametrize("agc_on", [True, False])
def test_frontend_matches_oracle(msdr, forc, agc_on):
    """70 channels (partial last group), 90 blocks in ragged updates, AGC moving the gain up and down, state carried."""
    C, nb = 70, 90
    codes = fl.adc_stream(C, 128 * nb, seed=7)
    g = msdr.Frontend(C, agc_on=agc_on)
    o = forc.frontend(C, on=int(agc_on))
    for c in (0, 5, 69):
        g.preset(codes[c, 0], c, 1)
        o.preset(c, codes[c, 0])
    outs, b0 = [], 0
    for n in (1, 2, 30, 3, 27, 27):
        outs.append(g.update(codes[:, b0 * 128:(b0 + n) * 128]))
        b0 += n
    yg = np.concatenate(outs, axis=1)
    yo = o.run(codes)
    assert_same(yg, yo, f"frontend agc_on={agc_on}")
    for c in (0, 1, 33, 69):
        _state_equal(g.get_state(c), o.state(c))
    if agc_on:
        assert len({g.get_state(c).multiplier for c in range(C)}) > 10  # channels ended up at different gains


def test_frontend_extreme_codes(msdr, forc):
    """Full 16-bit codes (accumulator wraps), rails (output saturates through the gain), constant input, 40x gain."""
    rng = np.random.default_rng(9)
    nb = 64
    codes = np.stack([rng.integers(0, 65536, 128 * nb).astype(np.uint16),
                      np.tile(np.r_[np.zeros(128), np.full(128, 65535)], nb // 2).astype(np.uint16),
                      np.full(128 * nb, 2048, np.uint16),
                      fl.adc_stream(1, 128 * nb, seed=3, bits=16)[0],
                      (2048 + 3 * rng.standard_normal(128 * nb)).astype(np.uint16)])
    for start, mx in ((0.25, 40.0), (39.5, 40.0), (1.0, 1.0)):
        g = msdr.Frontend(len(codes), agc_start=start, agc_max=mx)
        o = forc.frontend(len(codes), start, mx)
        assert_same(g.update(codes), o.run(codes), f"extreme start={start}")
        for c in range(len(codes)):
            _state_equal(g.get_state(c), o.state(c))


def test_amp_gain_multiplier(msdr, forc):
    for gain in (0.25, 1.0, 0.999999, 1.5, 40.0, 3.3e-6, 0.0, -2.0, 1e9, -1e9, 0.1, 17.123):
        assert msdr.frontend.amp_gain_multiplier(gain) == forc.amp_multiplier(gain)


def test_frontend_state_roundtrip_and_errors(msdr, forc):
    codes = fl.adc_stream(2, 128 * 20, seed=4)
    a = msdr.Frontend(2)
    y1 = a.update(codes[:, :128 * 10])
    b = msdr.Frontend(2)
    for c in range(2):
        b.set_state(c, a.get_state(c))  # migrate a channel between objects
    assert_same(np.concatenate([y1, b.update(codes[:, 128 * 10:])], axis=1), forc.frontend(2).run(codes), "migrated state")
    with pytest.raises(msdr.MsdrError):
        a.preset(0, 1, 5)
    with pytest.raises(msdr.MsdrError):
        msdr.Frontend(0)


def test_frontend_feeds_the_chain(msdr, forc, orc, K):
    """ADC codes -> front end -> receive chain on the GPU == the same two stages on the CPU checkers."""
    modes = msdr.synth.mixed_modes(40)
    codes = fl.adc_stream(len(modes), 128 * 48, seed=12)
    fg, fo = msdr.Frontend(len(modes)), forc.frontend(len(modes))
    g, o = configure_pair(msdr, orc, K, modes)
    xg, xo = fg.update(codes), fo.run(codes)
    assert_same(xg, xo, "front end")
    assert_same(g.update(xg), o.run(xo)[0], "front end + chain")


def test_dac_codes_and_amplifier_ops(msdr, forc):
    """msdr_op_dac_codes = ((s) + 32768) >> 4 (output_dac.cpp:143) on every int16 value; msdr_op_amplifier per-row multipliers."""
    L = msdr.capi.lib()
    s = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16).reshape(8, 8192)
    out = np.zeros(s.shape, np.uint16)
    assert L.msdr_op_dac_codes(0, msdr.capi.ptr(s), msdr.capi.ptr(out), 8, 8192, 8192) == 0
    assert np.array_equal(out, ((s.astype(np.int32) + 32768) >> 4).astype(np.uint16))
    assert out.min() == 0 and out.max() == 4095
    mults = np.array([forc.amp_multiplier(g) for g in (0.25, 1.0, 0.0, 1.7, -0.33, 40.0, 1e9, 0.999999)], np.int32)
    d = s.copy()
    assert L.msdr_op_amplifier(0, msdr.capi.ptr(mults), msdr.capi.ptr(d), 8, 8192, 8192) == 0
    for r in range(8):
        assert np.array_equal(d[r], forc.amp_apply(s[r], int(mults[r]))[0]), r


def test_packed_launch_and_spare_sms_change_nothing(msdr, forc, orc, K):
    """The options behind the two-stream pipeline of bench.py --with-frontend: the front end packed into a few multi-warp CTAs
    (msdr_frontend_set_option "sms") and the chain kernel leaving SMs free ("spare_sms") give the same bytes and the same state."""
    C, nb = 300, 12   # 10 channel groups, the last one partial
    codes = fl.adc_stream(C, 128 * nb, seed=17)
    outs = []
    for sms in (0, 3, 4, 1):   # 1 -> 10 groups in one CTA of 10 warps
        g = msdr.Frontend(C)
        g.set_option("sms", sms)
        outs.append((g.update(codes[:, :128 * 5]), g.update(codes[:, 128 * 5:]), [g.get_state(c).multiplier for c in (0, 31, 32, 299)]))
    o = forc.frontend(C)
    yo = o.run(codes)
    for a, b, st in outs:
        assert_same(np.concatenate([a, b], axis=1), yo, "packed front end")
        assert st == outs[0][2]
    modes = msdr.synth.mixed_modes(4096)
    x = msdr.synth.batch(modes[:64], 128 * 6)
    ys = []
    for spare in (0, 20, 147):
        gch, och = configure_pair(msdr, orc, K, modes[:64])
        gch.set_option("spare_sms", spare)
        ys.append(gch.update(x))
    assert_same(ys[0], och.run(x)[0], "spare_sms 0")
    assert np.array_equal(ys[0], ys[1]) and np.array_equal(ys[0], ys[2])
