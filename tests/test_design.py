"""Host-side designers (minimal-sdr_b200/design.py) against the reference's own designers compiled into oracle/_ref
and against the constants extracted from the sketch."""
import numpy as np
import pytest


def test_am_table_matches_committed_constants(msdr, K):
    am = msdr.design.calc_FIR_coeffs(102, 2800, 70, 0, 0.0, 24000)
    ref = np.array(K["FIR_AM_coeffs_bw2800_fs24000"])
    assert np.abs(am - ref).max() <= 1
    assert int(ref.sum()) == 32761 and ref[51] == 7645 and ref[50] == ref[52] == 6970  # SURVEY.md 8c probe values


@pytest.mark.parametrize("bw", [100, 1000, 2800, 3600, 5000])
@pytest.mark.parametrize("ftype", [0, 1])
def test_fir_designer_vs_reference(msdr, ref, bw, ftype):
    a = msdr.design.calc_FIR_coeffs(102, bw, 70, ftype, 0.0, 24000)
    b = ref.calc_fir_coeffs(102, float(bw), 70.0, ftype, 0.0, 24000.0)
    assert np.abs(a.astype(np.int32) - b).max() <= 1, (bw, ftype)


def test_biquad_designers_vs_reference(msdr, ref, K):
    d = msdr.design
    fs = K["AUDIO_SAMPLE_RATE_EXACT"]
    for f, q in [(9926.47, 0.54), (5514.7, 15.0), (300.0, 0.7071), (12000.0, 2.0)]:
        assert np.array_equal(d.biquad_lowpass(f, q, fs), ref.biquad_design(0, f, q)), (f, q)
        assert np.array_equal(d.biquad_highpass(f, q, fs), ref.biquad_design(1, f, q)), (f, q)
        assert np.array_equal(d.biquad_bandpass(f, q, fs), ref.biquad_design(2, f, q)), (f, q)
        assert np.array_equal(d.biquad_notch(f, q, fs), ref.biquad_design(3, f, q)), (f, q)
    for f, g, s in [(1000.0, -6.0, 1.0), (4000.0, -9.0, 0.7)]:  # boosts overflow Q2.30 (UB in the reference)
        assert np.array_equal(d.biquad_lowshelf(f, g, s, fs), ref.biquad_design(4, f, g, s)), (f, g, s)
        assert np.array_equal(d.biquad_highshelf(f, g, s, fs), ref.biquad_design(5, f, g, s)), (f, g, s)


def test_live_cascade_constants(msdr, K):
    corr = K["AUDIO_SAMPLE_RATE_EXACT"] / K["SAMPLE_RATE"]
    assert list(msdr.design.biquad_lowpass(K["IF"] * 0.9 * corr, 0.54)) == K["biquad1_lowpass_coef"]   # .ino:391-393
    assert list(msdr.design.biquad_notch(K["SAMPLE_RATE"] / 8 * corr, 15.0)) == K["biquad2_notch_coef"]  # .ino:356


def test_ssb_q_table_is_reversed_i(K):
    assert K["FIR_SSB_Q_coeffs"] == K["FIR_SSB_I_coeffs"][::-1]
    assert K["FIR_CW_Q_coeffs"] == K["FIR_CW_I_coeffs"][::-1]
    assert K["mode_enum"] == {"SYNCAM": 0, "AM": 1, "LSB": 2, "USB": 3, "CW": 4}
