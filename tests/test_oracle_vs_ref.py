"""Pin the C oracle (oracle/msdr_oracle.c) against the REFERENCE's own compiled sources (oracle/_ref).
Runs where /root/reference was available at build time or the prebuilt _ref library travelled."""
import numpy as np
import pytest

import oracle_lib as ol
from conftest import adversarial_inputs, wrap_coeffs


def test_mix_matches_reference(orc, ref):
    rng = np.random.default_rng(1)
    for x in adversarial_inputs(128 * 4, rng).values():
        for a, b in zip(orc.mix_fs4(x), ref.mix_fs4(x)):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("T", [4, 6, 86, 102, 256])
def test_fir_matches_reference(orc, ref, T):
    rng = np.random.default_rng(T)
    for trial in range(3):
        c = wrap_coeffs(T, rng) if trial else (rng.integers(-2000, 2000, T).astype(np.int16))
        for name, x in adversarial_inputs(128 * 5, rng).items():
            ya, ha, sa = orc.fir(c, x, hist_out=True)
            yb, hb, sb = ref.fir(c, x, hist_out=True)
            assert sa == sb == 0
            assert np.array_equal(ya, yb), (T, name)
            assert np.array_equal(ha, hb), (T, name)


def test_fir_block_partition_and_tails(orc, ref):
    """blockSize % 4 tail path (arm_fir_fast_q15.c:260-294) and arbitrary partitioning give the same stream."""
    rng = np.random.default_rng(7)
    c = wrap_coeffs(86, rng)
    x = rng.integers(-32768, 32768, 1000, dtype=np.int16)
    base = ref.fir(c, x, block=128)
    for blk in (1, 3, 5, 127, 128, 130, 1000):
        assert np.array_equal(ref.fir(c, x, block=blk), base), blk
        assert np.array_equal(orc.fir(c, x, block=blk), base), blk


def test_fir_odd_taps_status(orc, ref):
    assert orc.fir_init_status(85) == ref.fir_init_status(85) == -1  # ARM_MATH_ARGUMENT_ERROR
    assert orc.fir_init_status(86) == ref.fir_init_status(86) == 0


def test_fir_inplace_coefficient_rewrite(orc, ref):
    rng = np.random.default_rng(9)
    c1, c2 = wrap_coeffs(102, rng), wrap_coeffs(102, rng)
    x = rng.integers(-32768, 32768, 128 * 6, dtype=np.int16)
    assert np.array_equal(orc.fir(c1, x, recoef=(256, c2)), ref.fir(c1, x, recoef=(256, c2)))


def test_demod_and_sqrt_match_reference(orc, ref):
    rng = np.random.default_rng(3)
    I = rng.integers(-32768, 32768, 50000, dtype=np.int16)
    Q = rng.integers(-32768, 32768, 50000, dtype=np.int16)
    I[:4], Q[:4] = [32767, -32768, 100, 10000], [32767, -32768, 0, 10000]
    for kind in range(4):
        assert np.array_equal(orc.demod(kind, I, Q), ref.demod(kind, I, Q)), kind
    # corner values quoted in SURVEY.md A5b/A5c
    assert orc.demod(2, I[:1], Q[:1])[0] == -19197
    assert list(orc.demod(3, I[2:4], Q[2:4])) == [70, 9999]
    for v in list(rng.integers(1, 2 ** 31, 20000)) + [0, -5, 1, 2, 3, 2 ** 31 - 1, -2 ** 31]:
        assert orc.sqrt_q31(int(v)) == ref.sqrt_q31(int(v)), v


def test_biquad_matches_reference(orc, ref, K):
    rng = np.random.default_rng(5)
    x = rng.integers(-32768, 32768, 128 * 40, dtype=np.int16)
    lp, notch = K["biquad1_lowpass_coef"], K["biquad2_notch_coef"]
    hot = [int(1.9 * 2 ** 30), int(-1.7 * 2 ** 30), int(1.9 * 2 ** 30), int(-1.2 * 2 ** 30), int(0.5 * 2 ** 30)]
    rnd = [int(v) for v in rng.integers(-2 ** 31, 2 ** 31, 5)]
    for stages in ([(0, lp)], [(0, notch)], [(0, hot)], [(0, rnd)], [(0, lp), (1, notch)], [(0, lp), (1, notch), (2, hot), (3, lp)],
                   [(0, lp), (2, notch)], [(5, lp)], []):
        ya, da = orc.biquad(stages, x, definition_out=True)
        yb, db = ref.biquad(stages, x, definition_out=True)
        assert np.array_equal(ya, yb), stages
        assert np.array_equal(da, db), stages


def test_biquad_recoefficient_keeps_history(orc, ref, K):
    """setCoefficients mid-stream keeps x/y history and clears the residual (filter_biquad.cpp:95-98)."""
    rng = np.random.default_rng(6)
    x = rng.integers(-20000, 20000, 128 * 4, dtype=np.int16)
    outs = []
    for L in (orc, ref):
        y1, d1 = L.biquad([(0, K["biquad1_lowpass_coef"])], x[:256], definition_out=True)
        y2, d2 = L.biquad([(0, K["biquad2_notch_coef"])], x[256:], definition_out=True, definition_in=d1)
        outs.append((y1, y2, d2))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("am_q31", [False, True])
def test_chain_matches_reference(orc, ref, K, am_q31):
    import minimal_sdr_b200 as m
    rng = np.random.default_rng(11)
    modes = [ol.MODE_AM, ol.MODE_USB, ol.MODE_LSB, ol.MODE_CW, ol.MODE_AM, ol.MODE_USB]
    x = m.synth.batch(modes, 128 * 12)
    x[4] = rng.integers(-32768, 32768, x.shape[1], dtype=np.int16)
    x[5] = -32768
    ys = []
    for L in (orc, ref):
        ch = L.chain(len(modes), am_q31)
        for c, md in enumerate(modes):
            ch.set_mode(c, 1, md)
            if md in (ol.MODE_USB, ol.MODE_LSB):
                ch.fir_init(c, 1, K["FIR_SSB_I_coeffs"], K["FIR_SSB_Q_coeffs"])
            elif md == ol.MODE_CW:
                ch.fir_init(c, 1, K["FIR_CW_I_coeffs"], K["FIR_CW_Q_coeffs"])
            else:
                ch.fir_init(c, 1, K["FIR_AM_coeffs_bw2800_fs24000"], K["FIR_AM_coeffs_bw2800_fs24000"])
        ch.biquad_set_coefficients(0, 0, len(modes), 0, K["biquad1_lowpass_coef"])
        ch.biquad_set_coefficients(1, 0, len(modes), 0, K["biquad2_notch_coef"])
        y1, _ = ch.run(x[:, :128 * 5])
        y2, _ = ch.run(np.ascontiguousarray(x[:, 128 * 5:]))
        ys.append(np.concatenate([y1, y2], axis=1))
        ch.close()
    assert np.array_equal(ys[0], ys[1])
    assert np.abs(ys[0][:4].astype(np.int32)).max() > 100  # the chain actually passes signal


def test_reference_builds_agree_across_optimisation_levels(orc, K):
    """oracle/Makefile `ref_variants`: the reference sources at -O2 (default), -O2 -fno-tree-vectorize and -O3 -march=x86-64-v3 give
    the same chain output as the oracle on wrap / saturation inputs, so the faster build is a valid CPU baseline (bench.py)."""
    import os
    from chain_helpers import tables_for
    rng = np.random.default_rng(77)
    modes = [ol.MODE_AM, ol.MODE_USB, ol.MODE_LSB, ol.MODE_CW, ol.MODE_AM, ol.MODE_USB]
    x = np.stack([v for v in list(adversarial_inputs(128 * 6, rng).values())[:6]])
    tabs = {1: (wrap_coeffs(86, rng), wrap_coeffs(86, rng)), 4: (wrap_coeffs(102, rng), wrap_coeffs(102, rng))}

    def run(lib):
        o = lib.chain(len(modes))
        for c, md in enumerate(modes):
            o.set_mode(c, 1, md)
            assert o.fir_init(c, 1, *(tabs.get(c) or tables_for(K, md))) == 0
        o.biquad_set_coefficients(0, 0, len(modes), 0, K["biquad1_lowpass_coef"])
        o.biquad_set_coefficients(1, 0, len(modes), 0, K["biquad2_notch_coef"])
        y = o.run(x)[0]
        o.close()
        return y

    want = run(orc)
    seen = 0
    for name, path in ol.REF_VARIANTS.items():
        if not os.path.exists(path):
            continue
        assert np.array_equal(run(ol.CheckerLib("ref", path=path)), want), name
        seen += 1
    if seen == 0:
        pytest.skip("no oracle/_ref build here")


def test_freq_conv_matches_reference_class(orc, ref):
    """A6: orc_freq_conv == the reference's own freq_conv.cpp compiled where it lies, with arm_mult/add/sub_q15 built from the
    vendored portable clip_q31_to_q15 / __QADD16 / __QSUB16 (arm_math.h:555-560,721-765): every pairing of the q15 corner values
    (sum / difference saturation, -32768 * -32768), both directions, both pass settings."""
    if "freq_conv" not in ref._f:
        pytest.skip("oracle/_ref predates the freq_conv build")
    sys_path_golden = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", __import__("os").path.join(sys_path_golden, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    I, Q, oI, oQ = mg.freq_conv_inputs()
    for d in (0, 1):
        for ps in (0, 1):
            a, b = orc.freq_conv(d, ps, I, Q, oI, oQ), ref.freq_conv(d, ps, I, Q, oI, oQ)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (d, ps)
    # silent-drop paths of the class (freq_conv.cpp:40-47,64,111): nothing is transmitted
    assert ref.freq_conv_ex(0, 1, I[:128], Q[:128], oI[:128], oQ[:128], have_I=0)[2] is False
    assert ref.freq_conv_ex(0, 1, I[:128], Q[:128], oI[:128], oQ[:128], have_Q=0)[2] is False
    assert ref.freq_conv_ex(0, 1, I[:128], Q[:128], oI[:128], oQ[:128], fail_alloc=1)[2] is False
    assert ref.freq_conv_ex(0, 1, I[:128], Q[:128], oI[:128], oQ[:128])[2] is True


def test_vendored_saturating_primitives(ref):
    """The saturating q15 primitives the reference vendors (portable C, arm_math.h:555-560,721-765), on their corners."""
    if "qadd16" not in ref._f:
        pytest.skip("oracle/_ref predates the freq_conv build")
    f = ref._f
    pk = lambda hi, lo: ((hi & 0xFFFF) << 16) | (lo & 0xFFFF)
    assert f["qadd16"](pk(32767, -32768), pk(1, -1)) == pk(32767, -32768)      # both lanes saturate
    assert f["qadd16"](pk(-1, 5), pk(1, -5)) == pk(0, 0)
    assert f["qsub16"](pk(-32768, 32767), pk(1, -1)) == pk(-32768, 32767)
    assert f["qsub16"](pk(0, 0), pk(-32768, -32768)) == pk(32767, 32767)      # 0 - (-32768) saturates
    assert f["clip_q31_to_q15"]((-32768 * -32768) >> 15) == 32767              # the one q15 product that overflows
    assert f["clip_q31_to_q15"](-32769) == -32768 and f["clip_q31_to_q15"](12345) == 12345
