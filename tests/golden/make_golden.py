#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ from the REFERENCE's own code (oracle/_ref/libmsdr_ref.so, built from
/root/reference by oracle/Makefile).  The reference ships no test vectors; these are outputs of its compiled sources on
seeded inputs.  Run in the build container:  python tests/golden/make_golden.py
Fixtures are small .npz files; tests/test_oracle_golden.py checks the C oracle against them on any machine and
tests/test_gpu_chain.py checks the CUDA path against them on the B200 box (where /root/reference does not exist).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import oracle_lib as ol  # noqa: E402
from conftest import adversarial_inputs, wrap_coeffs  # noqa: E402
import minimal_sdr_b200 as m  # noqa: E402


def freq_conv_inputs():
    """A6 known-answer inputs: every pairing of the q15 corner values (saturation of the products' sum / difference, the
    -32768 * -32768 product that overflows q15) plus seeded full-range noise, against fs/4 tables and full-scale noise tables."""
    rng = np.random.default_rng(606)
    corners = np.array([-32768, -32767, -16384, -1, 0, 1, 16384, 32766, 32767], np.int16)
    g = np.array(np.meshgrid(corners, corners, corners, corners, indexing="ij")).reshape(4, -1)  # I, Q, oscI, oscQ: 6561 columns
    n = (g.shape[1] + 127) // 128 * 128 + 128 * 8
    I, Q, oI, oQ = (rng.integers(-32768, 32768, n, dtype=np.int16) for _ in range(4))
    I[:g.shape[1]], Q[:g.shape[1]], oI[:g.shape[1]], oQ[:g.shape[1]] = g
    k = np.arange(128 * 4)
    oI[-512:] = np.array([0, 32767, 0, -32767], np.int16)[k % 4]   # the fs/4 oscillator a user of the class would supply
    oQ[-512:] = np.array([32767, 0, -32767, 0], np.int16)[k % 4]
    return I, Q, oI, oQ


def freq_conv_kat(ref):
    I, Q, oI, oQ = freq_conv_inputs()
    kat = {"I": I, "Q": Q, "oscI": oI, "oscQ": oQ}
    for d in (0, 1):
        for ps in (0, 1):
            kat[f"I_dir{d}_pass{ps}"], kat[f"Q_dir{d}_pass{ps}"] = ref.freq_conv(d, ps, I, Q, oI, oQ)
    np.savez_compressed(os.path.join(HERE, "freqconv_kat.npz"), **kat)


def main():
    ref = ol.CheckerLib("ref")
    freq_conv_kat(ref)
    if "--only-freqconv" in sys.argv:
        return
    K = m.load_ref_constants()
    rng = np.random.default_rng(20261017)
    tabs = {
        "ssb_i": np.array(K["FIR_SSB_I_coeffs"], np.int16), "ssb_q": np.array(K["FIR_SSB_Q_coeffs"], np.int16),
        "cw_i": np.array(K["FIR_CW_I_coeffs"], np.int16), "cw_q": np.array(K["FIR_CW_Q_coeffs"], np.int16),
        "am": np.array(K["FIR_AM_coeffs_bw2800_fs24000"], np.int16),
        "wrap86": wrap_coeffs(86, rng), "wrap102": wrap_coeffs(102, rng), "wrap256": wrap_coeffs(256, rng),
        "tiny4": np.array([32767, -32768, 12345, -1], np.int16),
    }

    # ---- FIR known answers: every table on uniform + adversarial inputs, 6 blocks
    n = 128 * 6
    fir = {"tab_" + k: v for k, v in tabs.items()}
    inputs = adversarial_inputs(n, rng)
    for iname, x in inputs.items():
        fir["in_" + iname] = x
        for tname, c in tabs.items():
            fir[f"out_{tname}__{iname}"] = ref.fir(c, x)
    np.savez_compressed(os.path.join(HERE, "fir_kat.npz"), **fir)

    # ---- demod + sqrt corners
    I = rng.integers(-32768, 32768, 2048, dtype=np.int16)
    Q = rng.integers(-32768, 32768, 2048, dtype=np.int16)
    corners = np.array([[32767, 32767], [-32768, -32768], [100, 0], [10000, 10000], [0, 0], [-32768, 0], [0, -32768],
                        [32767, -32768], [1, 1], [-1, 1], [181, 181], [23170, 23170], [23171, 23170]], np.int16)
    I[:len(corners)], Q[:len(corners)] = corners[:, 0], corners[:, 1]
    dem = {"I": I, "Q": Q}
    for kind in range(4):
        dem[f"out{kind}"] = ref.demod(kind, I, Q)
    sq_in = np.concatenate([np.array([0, -1, 1, 2, 3, 4, 100, 10000, 200000000, 2 ** 31 - 1, -2 ** 31, 2 ** 30, 2 ** 30 - 1, 2 ** 29, 65536, 65535],
                                     np.int64), rng.integers(1, 2 ** 31, 4000)]).astype(np.int32)
    sq = np.array([ref.sqrt_q31(int(v)) for v in sq_in], np.int64)
    dem["sqrt_in"], dem["sqrt_out"], dem["sqrt_status"] = sq_in, sq[:, 0].astype(np.int32), sq[:, 1].astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "demod_kat.npz"), **dem)

    # ---- biquad: the live cascade coefficients, a saturating high-gain set, a 4-stage cascade
    lp, notch = np.array(K["biquad1_lowpass_coef"], np.int32), np.array(K["biquad2_notch_coef"], np.int32)
    hot = np.array([int(1.9 * 2 ** 30), int(-1.7 * 2 ** 30), int(1.9 * 2 ** 30), int(-1.2 * 2 ** 30), int(0.5 * 2 ** 30)], np.int32)
    x = rng.integers(-32768, 32768, 128 * 8, dtype=np.int16)
    bq = {"x": x, "lp": lp, "notch": notch, "hot": hot}
    for name, stages in {"lp": [(0, lp)], "notch": [(0, notch)], "hot": [(0, hot)], "lp_notch_hot_lp": [(0, lp), (1, notch), (2, hot), (3, lp)],
                         "gap": [(0, lp), (2, notch)]}.items():
        y, d = ref.biquad(stages, x, definition_out=True)
        bq["y_" + name], bq["def_" + name] = y, d
    np.savez_compressed(os.path.join(HERE, "biquad_kat.npz"), **bq)

    # ---- whole chain: 12 channels x 10 blocks, modes cycling, reference tables, live biquads; plus a q31 (Teensy 3.2) run
    C, nb = 12, 10
    modes = [ol.MODE_AM, ol.MODE_USB, ol.MODE_LSB, ol.MODE_CW] * 3
    xin = m.synth.batch(modes, nb * 128)
    xin[8] = rng.integers(-32768, 32768, nb * 128, dtype=np.int16)  # full-scale noise through USB? (mode AM at index 8)
    xin[9] = -32768
    chain = {"modes": np.array(modes, np.int32), "x": xin}
    for q31 in (0, 1):
        ch = ref.chain(C, am_q31=bool(q31))
        for c, md in enumerate(modes):
            ch.set_mode(c, 1, md)
            if md in (ol.MODE_USB, ol.MODE_LSB):
                ch.fir_init(c, 1, tabs["ssb_i"], tabs["ssb_q"])
            elif md == ol.MODE_CW:
                ch.fir_init(c, 1, tabs["cw_i"], tabs["cw_q"])
            else:
                ch.fir_init(c, 1, tabs["am"], tabs["am"])
        ch.biquad_set_coefficients(0, 0, C, 0, lp)
        ch.biquad_set_coefficients(1, 0, C, 0, notch)
        y, _ = ch.run(xin)
        chain[f"y_q31_{q31}"] = y
        ch.close()
    np.savez_compressed(os.path.join(HERE, "chain_kat.npz"), **chain)

    # front-end conditioning: the reference's DC-blocking loop, AudioAmplifier and AGC()
    import frontend_lib as fl
    rf = fl.Ref()
    codes = np.concatenate([fl.adc_stream(1, 128 * 12, seed=11)[0], rng.integers(0, 65536, 128 * 4).astype(np.uint16),
                            np.full(128 * 2, 65535, np.uint16), np.zeros(128 * 2, np.uint16)])
    st0 = (int(codes[0]) << 14, 0)
    hout, x1, y1 = rf.hpf(codes, *st0)
    fe = {"hpf_in": codes, "hpf_state": np.array(st0, np.int64), "hpf_out": hout, "hpf_state_out": np.array([x1, y1], np.int64)}
    blk = rng.integers(-32768, 32768, 128).astype(np.int16)
    blk[:4] = [-32768, 32767, 0, -1]
    gains = np.array([0.25, 0.9, 1.5, 40.0, 17.123, -2.0, 1e9], np.float32)
    fe["amp_in"], fe["amp_gains"] = blk, gains
    fe["amp_mults"] = np.array([rf.amp_multiplier(g) for g in gains], np.int64)
    fe["amp_out"] = np.stack([rf.amp_block(g, blk)[0] for g in gains])
    levels = np.abs(rng.normal(0, 1, 300)) * rng.choice([30, 300, 3000, 12000, 16000, 20000, 30000, 40000], 300)
    ablocks = (rng.normal(0, 1, (300, 128)) * levels[:, None] * 0.4).clip(-32768, 32767).astype(np.int16)
    ablocks[3] = 0
    ablocks[4, 9] = -32768
    fe["agc_blocks"] = ablocks
    fe["agc_val"], fe["agc_mult"] = fl.Ref.agc_trajectory(ablocks, 0.25, 40.0)
    np.savez_compressed(os.path.join(HERE, "frontend_kat.npz"), **fe)

    # LMS notch / noise reduction: the reference's own block, one stream per fresh process
    import anr_lib as al
    ax = al.audio_stream(2, 128 * 30, seed=31)
    anr = {"x": ax}
    for mode in (1, 2):
        anr[f"y_mode{mode}"] = np.stack([al.ref_anr_run(mode, ax[c]) for c in range(2)])
    np.savez_compressed(os.path.join(HERE, "anr_kat.npz"), **anr)

    # synchronous-AM PLL: the reference's own `case SYNCAM` arm, one stream per fresh process
    import syncam_lib as sl
    sI, sQ = sl.baseband(2, 128 * 40, seed=41)
    np.savez_compressed(os.path.join(HERE, "syncam_kat.npz"), I=sI, Q=sQ, y=np.stack([sl.ref_syncam_run(sI[c], sQ[c]) for c in range(2)]))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
