"""The C oracle against the committed golden vectors (tests/golden/*.npz, produced from the reference's own compiled
sources by tests/golden/make_golden.py).  Runs anywhere — no /root/reference, no GPU."""
import os

import numpy as np

import oracle_lib as ol

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_fir_golden(orc):
    z = np.load(os.path.join(G, "fir_kat.npz"))
    tabs = {k[4:]: z[k] for k in z.files if k.startswith("tab_")}
    ins = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    n = 0
    for t, c in tabs.items():
        for i, x in ins.items():
            assert np.array_equal(orc.fir(c, x), z[f"out_{t}__{i}"]), (t, i)
            n += 1
    assert n == len(tabs) * len(ins) >= 60


def test_fir_restated_direct_form(orc):
    """y[n] = ssat16((wrap32(sum_k c[k] u[n-(T-1)+k])) >> 15) with zero initial history."""
    z = np.load(os.path.join(G, "fir_kat.npz"))
    c, x = z["tab_wrap86"].astype(np.int64), z["in_uniform"].astype(np.int64)
    T = len(c)
    u = np.concatenate([np.zeros(T - 1, np.int64), x])
    acc = np.array([np.dot(c, u[n:n + T]) for n in range(len(x))])
    acc = ((acc + 2 ** 31) % 2 ** 32) - 2 ** 31
    y = np.clip(acc >> 15, -32768, 32767).astype(np.int16)
    assert np.array_equal(y, z["out_wrap86__uniform"])


def test_demod_golden(orc):
    z = np.load(os.path.join(G, "demod_kat.npz"))
    for kind in range(4):
        assert np.array_equal(orc.demod(kind, z["I"], z["Q"]), z[f"out{kind}"]), kind
    for v, o, s in zip(z["sqrt_in"], z["sqrt_out"], z["sqrt_status"]):
        assert orc.sqrt_q31(int(v)) == (int(o), int(s))


def test_biquad_golden(orc):
    z = np.load(os.path.join(G, "biquad_kat.npz"))
    cases = {"lp": [(0, z["lp"])], "notch": [(0, z["notch"])], "hot": [(0, z["hot"])],
             "lp_notch_hot_lp": [(0, z["lp"]), (1, z["notch"]), (2, z["hot"]), (3, z["lp"])], "gap": [(0, z["lp"]), (2, z["notch"])]}
    for name, stages in cases.items():
        y, d = orc.biquad(stages, z["x"], definition_out=True)
        assert np.array_equal(y, z["y_" + name]), name
        assert np.array_equal(d, z["def_" + name]), name
    assert np.abs(z["y_hot"].astype(np.int32)).max() == 32768  # the saturating case really saturates


def test_chain_golden(orc):
    import minimal_sdr_b200 as m
    K = m.load_ref_constants()
    z = np.load(os.path.join(G, "chain_kat.npz"))
    modes, x = z["modes"], z["x"]
    for q31 in (0, 1):
        ch = orc.chain(len(modes), bool(q31))
        for c, md in enumerate(modes):
            ch.set_mode(c, 1, int(md))
            if md in (ol.MODE_USB, ol.MODE_LSB):
                ch.fir_init(c, 1, K["FIR_SSB_I_coeffs"], K["FIR_SSB_Q_coeffs"])
            elif md == ol.MODE_CW:
                ch.fir_init(c, 1, K["FIR_CW_I_coeffs"], K["FIR_CW_Q_coeffs"])
            else:
                ch.fir_init(c, 1, K["FIR_AM_coeffs_bw2800_fs24000"], K["FIR_AM_coeffs_bw2800_fs24000"])
        ch.biquad_set_coefficients(0, 0, len(modes), 0, K["biquad1_lowpass_coef"])
        ch.biquad_set_coefficients(1, 0, len(modes), 0, K["biquad2_notch_coef"])
        y, _ = ch.run(x)
        assert np.array_equal(y, z[f"y_q31_{q31}"]), q31
        ch.close()


def test_freq_conv_semantics(orc):
    """A6 (parity unpinned, see oracle header): fs/4 tables, pass flag polarity, direction symmetry."""
    rng = np.random.default_rng(2)
    I = rng.integers(-32768, 32768, 128, dtype=np.int16)
    Q = rng.integers(-32768, 32768, 128, dtype=np.int16)
    k = np.arange(128)
    oscI = np.array([0, 32767, 0, -32767], np.int16)[k % 4]
    oscQ = np.array([32767, 0, -32767, 0], np.int16)[k % 4]
    i0, q0 = orc.freq_conv(0, 0, I, Q, oscI, oscQ)  # pass == 0 forwards unchanged (freq_conv.cpp:49-56)
    assert np.array_equal(i0, I) and np.array_equal(q0, Q)
    i1, q1 = orc.freq_conv(0, 1, I, Q, oscI, oscQ)
    a = lambda u, v: np.clip((u.astype(np.int32) * v) >> 15, -32768, 32767)
    assert np.array_equal(i1, np.clip(a(I, oscQ) + a(Q, oscI), -32768, 32767).astype(np.int16))
    assert np.array_equal(q1, np.clip(a(Q, oscQ) - a(I, oscI), -32768, 32767).astype(np.int16))
    i2, q2 = orc.freq_conv(1, 1, I, Q, oscI, oscQ)
    assert np.array_equal(q2, np.clip(a(Q, oscQ) + a(I, oscI), -32768, 32767).astype(np.int16))
    assert np.array_equal(i2, np.clip(a(I, oscQ) - a(Q, oscI), -32768, 32767).astype(np.int16))


def test_freq_conv_golden(orc):
    """A6 known answers generated from the reference's own freq_conv.cpp (tests/golden/make_golden.py::freq_conv_kat)."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "freqconv_kat.npz"))
    for d in (0, 1):
        for ps in (0, 1):
            i, q = orc.freq_conv(d, ps, z["I"], z["Q"], z["oscI"], z["oscQ"])
            assert np.array_equal(i, z[f"I_dir{d}_pass{ps}"]) and np.array_equal(q, z[f"Q_dir{d}_pass{ps}"]), (d, ps)
