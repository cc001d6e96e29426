"""Per-stage known-answer tests: each CUDA stage operator (msdr_op_*, through the C ABI) against the CPU oracle and the
committed golden vectors.  Bit-exact."""
import os

import numpy as np
import pytest

from conftest import adversarial_inputs, wrap_coeffs

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rows(d):
    names = sorted(d)
    return names, np.stack([d[k] for k in names])


def test_mix(msdr, orc):
    rng = np.random.default_rng(1)
    names, x = _rows(adversarial_inputs(128 * 3, rng))
    I, Q = np.empty_like(x), np.empty_like(x)
    L, p = msdr.capi.lib(), msdr.capi.ptr
    assert L.msdr_op_mix_fs4(0, p(x), p(I), p(Q), x.shape[0], x.shape[1], x.shape[1]) == 0
    for r, nme in enumerate(names):
        a, b = orc.mix_fs4(x[r])
        assert np.array_equal(I[r], a) and np.array_equal(Q[r], b), nme


@pytest.mark.parametrize("T", [4, 6, 86, 102, 256, 510])
def test_fir_fast_q15(msdr, orc, T):
    rng = np.random.default_rng(T)
    L, p = msdr.capi.lib(), msdr.capi.ptr
    names, x = _rows(adversarial_inputs(128 * 9 + 4 * (T % 3), rng))
    for trial in range(2):
        c = wrap_coeffs(T, rng) if trial else rng.integers(-1500, 1500, T).astype(np.int16)
        hist = rng.integers(-32768, 32768, (x.shape[0], T - 1), dtype=np.int16) if trial else np.zeros((x.shape[0], T - 1), np.int16)
        h_io, y = hist.copy(), np.empty_like(x)
        assert L.msdr_op_fir_fast_q15(0, T, p(c), p(h_io), p(x), p(y), x.shape[0], x.shape[1], x.shape[1]) == 0
        for r, nme in enumerate(names):
            # oracle with a pre-loaded history: prepend it to the stream and drop the warm-up outputs
            full = orc.fir(c, np.concatenate([hist[r], x[r]]))
            assert np.array_equal(y[r], full[T - 1:]), (T, nme)
            assert np.array_equal(h_io[r], np.concatenate([hist[r], x[r]])[-(T - 1):]), (T, nme)


def test_fir_odd_taps_rejected(msdr):
    L, p = msdr.capi.lib(), msdr.capi.ptr
    x = np.zeros((1, 128), np.int16)
    assert L.msdr_op_fir_fast_q15(0, 85, p(x), None, p(x), p(x.copy()), 1, 128, 128) == msdr.capi.ERR_ARGUMENT


def test_fir_golden(msdr):
    z = np.load(os.path.join(G, "fir_kat.npz"))
    L, p = msdr.capi.lib(), msdr.capi.ptr
    ins = sorted(k[3:] for k in z.files if k.startswith("in_"))
    x = np.stack([z["in_" + k] for k in ins])
    for t in sorted(k[4:] for k in z.files if k.startswith("tab_")):
        c = np.ascontiguousarray(z["tab_" + t])
        y = np.empty_like(x)
        assert L.msdr_op_fir_fast_q15(0, len(c), p(c), None, p(x), p(y), x.shape[0], x.shape[1], x.shape[1]) == 0
        for r, i in enumerate(ins):
            assert np.array_equal(y[r], z[f"out_{t}__{i}"]), (t, i)


def test_demod_and_sqrt(msdr, orc):
    z = np.load(os.path.join(G, "demod_kat.npz"))
    L, p = msdr.capi.lib(), msdr.capi.ptr
    rng = np.random.default_rng(4)
    I = np.concatenate([z["I"], rng.integers(-32768, 32768, 1 << 16, dtype=np.int16)])
    Q = np.concatenate([z["Q"], rng.integers(-32768, 32768, 1 << 16, dtype=np.int16)])
    for kind in range(4):
        out = np.empty_like(I)
        assert L.msdr_op_demod(0, kind, p(I), p(Q), p(out), 1, I.size, I.size) == 0
        assert np.array_equal(out[:z["I"].size], z[f"out{kind}"]), kind
        assert np.array_equal(out, orc.demod(kind, I, Q)), kind
    # every (I, Q) with |I|,|Q| near the int16 limits, where I^2+Q^2 crosses float rounding boundaries / wraps
    edge = np.array([-32768, -32767, -23171, -23170, -1, 0, 1, 181, 182, 23170, 23171, 32766, 32767], np.int16)
    Ie, Qe = [a.ravel().copy() for a in np.meshgrid(edge, edge)]
    for kind in (2, 3):
        out = np.empty_like(Ie)
        assert L.msdr_op_demod(0, kind, p(Ie), p(Qe), p(out), 1, Ie.size, Ie.size) == 0
        assert np.array_equal(out, orc.demod(kind, Ie, Qe)), kind
    sin = np.ascontiguousarray(z["sqrt_in"])
    sout, sst = np.empty_like(sin), np.empty_like(sin)
    assert L.msdr_op_sqrt_q31(0, p(sin), p(sout), p(sst), sin.size) == 0
    assert np.array_equal(sout, z["sqrt_out"]) and np.array_equal(sst, z["sqrt_status"])


def test_sqrt_q31_dense_sweep(msdr, orc):
    """4 M inputs incl. every power of two +-1: the float-seeded Newton iteration must agree everywhere."""
    L, p = msdr.capi.lib(), msdr.capi.ptr
    rng = np.random.default_rng(8)
    v = np.concatenate([rng.integers(-2 ** 31, 2 ** 31, 1 << 18), np.array([(1 << k) + d for k in range(1, 31) for d in (-1, 0, 1)])]).astype(np.int32)
    out = np.empty_like(v)
    assert L.msdr_op_sqrt_q31(0, p(v), p(out), None, v.size) == 0
    exp = np.array([orc.sqrt_q31(int(t))[0] for t in v], np.int32)
    assert np.array_equal(out, exp)


def test_biquad(msdr, orc, K):
    zb = np.load(os.path.join(G, "biquad_kat.npz"))
    L, p = msdr.capi.lib(), msdr.capi.ptr
    rng = np.random.default_rng(5)
    lp, notch, hot = zb["lp"], zb["notch"], zb["hot"]
    cases = {"lp": [(0, lp)], "notch": [(0, notch)], "hot": [(0, hot)], "lp_notch_hot_lp": [(0, lp), (1, notch), (2, hot), (3, lp)],
             "gap": [(0, lp), (2, notch)], "rand": [(0, rng.integers(-2 ** 31, 2 ** 31, 5).astype(np.int32))], "none": []}
    names = sorted(cases)
    x0 = zb["x"]
    defs = np.zeros((len(names), 32), np.int32)
    for r, nme in enumerate(names):  # definition[] as setCoefficients leaves it, taken from the oracle's implementation
        _, d = orc.biquad(cases[nme], np.zeros(0, np.int16), definition_out=True)
        defs[r] = d
    data = np.stack([x0] * len(names))
    d_io = defs.copy()
    assert L.msdr_op_biquad(0, p(d_io), p(data), len(names), x0.size, x0.size) == 0
    for r, nme in enumerate(names):
        y, d = orc.biquad(cases[nme], x0, definition_out=True)
        assert np.array_equal(data[r], y), nme
        assert np.array_equal(d_io[r], d), nme
        if "y_" + nme in zb.files:
            assert np.array_equal(data[r], zb["y_" + nme]) and np.array_equal(d_io[r], zb["def_" + nme]), nme
    # second call continues from the returned state
    data2 = np.stack([rng.integers(-32768, 32768, 256, dtype=np.int16)] * len(names))
    x2 = data2[0].copy()
    assert L.msdr_op_biquad(0, p(d_io), p(data2), len(names), 256, 256) == 0
    for r, nme in enumerate(names):
        y = orc.biquad(cases[nme], np.concatenate([x0, x2]))
        assert np.array_equal(data2[r], y[x0.size:]), nme


def test_freq_conv(msdr, orc):
    L, p = msdr.capi.lib(), msdr.capi.ptr
    rng = np.random.default_rng(2)
    I = rng.integers(-32768, 32768, (5, 128), dtype=np.int16)
    Q = rng.integers(-32768, 32768, (5, 128), dtype=np.int16)
    I[0], Q[0] = -32768, -32768
    osc = {"fs4": (np.array([0, 32767, 0, -32767], np.int16)[np.arange(128) % 4], np.array([32767, 0, -32767, 0], np.int16)[np.arange(128) % 4]),
           "rand": (rng.integers(-32768, 32768, 128, dtype=np.int16), rng.integers(-32768, 32768, 128, dtype=np.int16)),
           "min": (np.full(128, -32768, np.int16), np.full(128, -32768, np.int16))}
    for (oi, oq) in osc.values():
        for d in (0, 1):
            for ps in (0, 1):
                i2, q2 = I.copy(), Q.copy()
                assert L.msdr_op_freq_conv(0, d, ps, p(i2), p(q2), p(oi), p(oq), 5, 128, 128) == 0
                for r in range(5):
                    a, b = orc.freq_conv(d, ps, I[r], Q[r], oi, oq)
                    assert np.array_equal(i2[r], a) and np.array_equal(q2[r], b), (d, ps, r)


def test_freq_conv_reference_known_answers(msdr):
    """A6 against tests/golden/freqconv_kat.npz — outputs of the reference's own freq_conv.cpp (compiled with arm_mult/add/sub_q15
    built from the vendored saturating primitives, oracle/ref_q15_prims.c) on every pairing of the q15 corner values."""
    import os
    L, p = msdr.capi.lib(), msdr.capi.ptr
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "freqconv_kat.npz"))
    n = z["I"].size
    for d in (0, 1):
        for ps in (0, 1):
            i2, q2 = z["I"].copy(), z["Q"].copy()
            assert L.msdr_op_freq_conv(0, d, ps, p(i2), p(q2), p(z["oscI"]), p(z["oscQ"]), 1, n, n) == 0
            assert np.array_equal(i2, z[f"I_dir{d}_pass{ps}"]) and np.array_equal(q2, z[f"Q_dir{d}_pass{ps}"]), (d, ps)
