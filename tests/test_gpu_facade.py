"""The C++ façade (include/msdr/Audio.h): the sketch's object graph rebuilt on it (minimal-sdr_b200/host/sketch_port.cpp) and
stand-alone AudioFilterBiquad / AudioEffectFreqConv objects, against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from chain_helpers import assert_same, tables_for

pytestmark = pytest.mark.gpu
HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "minimal-sdr_b200", "host")


def _build():
    subprocess.run(["make", "-s", "-C", HOST], check=True)


@pytest.mark.parametrize("mode", [ol.MODE_AM, ol.MODE_USB, ol.MODE_LSB, ol.MODE_CW, ol.MODE_SYNCAM])
def test_sketch_port(msdr, orc, K, tmp_path, mode):
    _build()
    C, NB = 37, 11
    x = msdr.synth.batch([mode] * C, NB * 128)
    cI, cQ = tables_for(K, mode)
    lp, notch = np.array(K["biquad1_lowpass_coef"], np.int32), np.array(K["biquad2_notch_coef"], np.int32)
    with open(tmp_path / "tables.bin", "wb") as f:
        f.write(np.int32(len(cI)).tobytes() + cI.tobytes() + cQ.tobytes() + lp.tobytes() + notch.tobytes())
    x.tofile(tmp_path / "in.bin")
    r = subprocess.run([os.path.join(HOST, "sketch_port"), str(C), str(NB), str(mode), str(tmp_path / "tables.bin"), str(tmp_path / "in.bin"),
                        str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    y = np.fromfile(tmp_path / "out.bin", np.int16).reshape(C, NB * 128)
    o = orc.chain(C)
    o.set_mode(0, C, mode)
    o.fir_init(0, C, cI, cQ)
    o.biquad_set_coefficients(0, 0, C, 0, lp)
    o.biquad_set_coefficients(1, 0, C, 0, notch)
    if mode == ol.MODE_SYNCAM:  # the sketch's default mode: PLL demodulator, float path => stated tolerance (tests/test_gpu_syncam.py)
        e = o.run(x)[0].astype(np.float64)
        assert np.abs(y - e).max() <= 2 and np.sqrt(np.mean((y - e) ** 2)) <= 1e-5 * np.sqrt(np.mean(e ** 2)) + 0.02
    else:
        assert_same(y, o.run(x)[0], f"sketch_port mode {mode}")
    assert np.abs(y.astype(np.int32)).max() > 300


def test_standalone_objects(msdr, orc, tmp_path):
    _build()
    C, NB = 5, 7
    rng = np.random.default_rng(12)
    x = rng.integers(-32768, 32768, (C, NB * 128), dtype=np.int16)
    x[0] = -32768
    x.tofile(tmp_path / "in.bin")
    outs = [str(tmp_path / n) for n in ("ob.bin", "oi.bin", "oq.bin")]
    r = subprocess.run([os.path.join(HOST, "facade_objects"), str(C), str(NB), str(tmp_path / "in.bin")] + outs, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ob, oi, oq = [np.fromfile(p, np.int16).reshape(C, NB * 128) for p in outs]
    d = msdr.design
    lp, notch = d.biquad_lowpass(3000.0, 0.7071), d.biquad_notch(5514.7, 15.0)
    c2 = d.biquad_double_to_int([0.2, 0.4, 0.2, -0.5, 0.3])
    for c in range(C):
        y1, d1 = orc.biquad([(0, lp), (1, notch)], x[c, :3 * 128], definition_out=True)
        y2 = orc.biquad([(0, c2)], x[c, 3 * 128:], definition_in=d1)
        assert np.array_equal(ob[c], np.concatenate([y1, y2])), c
    k = np.arange(128)
    oscI = np.array([0, 32767, 0, -32767], np.int16)[k % 4]
    oscQ = np.array([32767, 0, -32767, 0], np.int16)[k % 4]
    for c in range(C):
        for b in range(NB):
            I = x[c, b * 128:(b + 1) * 128]
            Q = I[::-1]
            ei, eq = orc.freq_conv(1 if b >= 2 else 0, 0 if b == NB - 1 else 1, I, Q, oscI, oscQ)
            assert np.array_equal(oi[c, b * 128:(b + 1) * 128], ei) and np.array_equal(oq[c, b * 128:(b + 1) * 128], eq), (c, b)


def test_frontend_objects(msdr, tmp_path):
    """Frontend (DC block + amp_adc + AGC) fed one audio block per call, and a stand-alone AudioAmplifier in an AudioConnection
    graph, against the CPU checker (pinned to the reference's compiled code by tests/test_frontend_cpu.py)."""
    import frontend_lib as fl
    _build()
    C, NB = 6, 40
    codes = fl.adc_stream(C, NB * 128, seed=21)
    codes.tofile(tmp_path / "codes.bin")
    outs = [str(tmp_path / n) for n in ("fe.bin", "amp.bin", "agc.bin")]
    r = subprocess.run([os.path.join(HOST, "frontend_objects"), str(C), str(NB), str(tmp_path / "codes.bin")] + outs, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ofe = np.fromfile(outs[0], np.int16).reshape(C, NB * 128)
    oamp = np.fromfile(outs[1], np.int16).reshape(C, NB * 128)
    agc = np.fromfile(outs[2], np.float32)
    forc = fl.Orc()
    o = forc.frontend(C)
    for c in range(C):
        o.preset(c, codes[0, 0])  # Frontend::begin presets every channel with the same first reading
    assert np.array_equal(ofe, o.run(codes))
    assert np.array_equal(agc.view(np.uint32), np.array([o.state(c)["agc_val"] for c in range(C)], np.float32).view(np.uint32))
    x = codes.view(np.int16)
    for b in range(NB):
        gain = 1.7 if b < 3 else 1.0 if b == 3 else 0.0 if b == 4 else -0.33
        for c in range(C):
            exp, sent = forc.amp_apply(x[c, b * 128:(b + 1) * 128], forc.amp_multiplier(gain))
            assert np.array_equal(oamp[c, b * 128:(b + 1) * 128], exp), (c, b)


def test_fir_object_and_processor_usage(msdr, orc, K, tmp_path):
    """AudioFilterFIR::begin/end/update in an AudioConnection graph (arm_fir_fast_q15 per channel, FIR_PASSTHRU, end(), failed init) and
    AudioProcessorUsageMax()/AudioProcessorUsageMaxReset() fed by a Receiver (Minimal-SDR.ino:424-426); host/fir_usage_objects.cpp."""
    _build()
    C, NB = 6, 9
    rng = np.random.default_rng(33)
    x = rng.integers(-32768, 32768, (C, NB * 128), dtype=np.int16)
    x[1] = -32768
    taps = np.array(K["FIR_SSB_I_coeffs"], np.int16)
    with open(tmp_path / "taps.bin", "wb") as f:
        f.write(np.int32(taps.size).tobytes() + taps.tobytes())
    x.tofile(tmp_path / "in.bin")
    r = subprocess.run([os.path.join(HOST, "fir_usage_objects"), str(C), str(NB), str(tmp_path / "taps.bin"), str(tmp_path / "in.bin"),
                        str(tmp_path / "out.bin"), str(tmp_path / "sent.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    y = np.fromfile(tmp_path / "out.bin", np.int16).reshape(C, NB * 128)
    sent = np.fromfile(tmp_path / "sent.bin", np.uint8)
    assert sent.tolist() == [1, 1, 1, 0, 1, 0, 1, 1, 1]  # end() and the odd tap count transmit nothing; FIR_PASSTHRU forwards
    for c in range(C):
        assert np.array_equal(y[c, :3 * 128], orc.fir(taps, x[c, :3 * 128])), c
        assert np.array_equal(y[c, 4 * 128:5 * 128], x[c, 4 * 128:5 * 128]), c
        assert np.array_equal(y[c, 6 * 128:], orc.fir(taps, x[c, 6 * 128:])), c  # begin() zeroed the delay line
    assert "usage before 0.0000%" in r.stdout
