"""Host-side coefficient designers (control path).

* `calc_FIR_coeffs`  — Kaiser-windowed-sinc designer of the sketch (Minimal-SDR.ino:782-899: calc_FIR_coeffs, m_sinc,
  Izero), restated with the reference's float/double promotion rules (float32 variables, double literals).  libm
  `sinf`/`powf` differ between newlib (Teensy), glibc and numpy at the ulp level, so design output is pinned to the
  compiled reference only to +-1 LSB (tests/test_design.py); kernels are always tested with integer taps.
* `biquad_*` — the RBJ-cookbook designers of AudioFilterBiquad (filter_biquad.h:44-149): double math, C truncation
  to int, result = {b0,b1,b2,a1,a2} in Q2.30 as passed to setCoefficients(stage, const int*).
"""
import math

import numpy as np

f32 = np.float32
PI = 3.1415926535897932384626433832795  # Teensy core (double); arm_math.h:365 only defines PI if absent
PIH = PI / 2
AUDIO_SAMPLE_RATE_EXACT = 44117.64706  # Teensy core constant the Audio library hard-wires


def _izero(x):
    """Minimal-SDR.ino:880-897, float32 arithmetic."""
    x = f32(x)
    x2 = f32(float(x) / 2.0)
    summe, ds, di = f32(1.0), f32(1.0), f32(1.0)
    errorlimit = f32(1e-9)
    while True:
        tmp = f32(x2 / di)
        tmp = f32(tmp * tmp)
        ds = f32(ds * tmp)
        summe = f32(summe + ds)
        di = f32(float(di) + 1.0)
        if not (ds >= f32(errorlimit * summe)):
            break
    return summe


def _m_sinc(m, fc):
    """Minimal-SDR.ino:871-878."""
    if m == 0:
        return f32(1.0)
    x = f32(m * PIH)
    return f32(np.sin(f32(x * fc), dtype=f32) / f32(fc * x))


def _trunc16(v):
    """C float -> int16 narrowing as the Cortex-M4/x86 do it: truncate to int32, keep the low 16 bits."""
    return int(np.int16(np.int32(np.trunc(v)) & 0xFFFF if abs(float(v)) < 2 ** 31 else 0))


def calc_FIR_coeffs(numCoeffs, fc, Astop, ftype=0, dfc=0.0, Fsamprate=24000.0):
    """Returns int16[numCoeffs] like `calc_FIR_coeffs(coeffs, numCoeffs, fc, Astop, type, dfc, Fsamprate)`.
    Types: 0 low-pass, 1 high-pass, 2 band-pass, 3 notch (type 4, Hilbert, writes a different layout in the
    reference and is not used by the sketch)."""
    if ftype not in (0, 1, 2, 3):
        raise ValueError("type must be 0..3")
    fc = f32(f32(fc) / f32(Fsamprate))
    dfc = f32(f32(dfc) / f32(Fsamprate))
    Astop = f32(Astop)
    if Astop < f32(20.96):
        Beta = f32(0.0)
    elif Astop >= f32(50.0):
        Beta = f32(0.1102 * (float(Astop) - 8.71))
    else:
        Beta = f32(0.5842 * float(np.power(f32(float(Astop) - 20.96), f32(0.4), dtype=f32)) + 0.07886 * (float(Astop) - 20.96))
    izb = _izero(Beta)
    if ftype == 0:
        fcf, nc = f32(float(fc) * 2.0), numCoeffs
    elif ftype == 1:
        fcf, nc = f32(-fc), 2 * (numCoeffs // 2)
    else:
        fcf, nc = dfc, 2 * (numCoeffs // 2)
    n_written = max(numCoeffs, nc + 1)
    coeffs = np.zeros(n_written + 1, np.int16)
    jj = 0
    for ii in range(-nc, nc, 2):
        x = f32(f32(ii) / f32(nc))
        w = f32(_izero(f32(Beta * np.sqrt(f32(f32(1.0) - f32(x * x)), dtype=f32))) / izb)
        v = f32(f32(f32(fcf * _m_sinc(ii, fcf)) * w) * f32(32767))
        coeffs[jj] = _trunc16(v)
        jj += 1
    if ftype == 1:
        coeffs[nc // 2] += 1
    elif ftype in (2, 3):
        sgn = 2.0 if ftype == 2 else -2.0
        for j in range(nc + 1):
            g = f32(f32(sgn) * np.cos(f32(PIH * (2 * j - nc) * float(fc)), dtype=f32))
            coeffs[j] = _trunc16(f32(f32(coeffs[j]) * g))
        if ftype == 3:
            coeffs[nc // 2] += 1
    return coeffs[:numCoeffs].copy()


# ---- AudioFilterBiquad designers (filter_biquad.h) --------------------------------------------------------

def _c_int(v):
    """C double -> int conversion: truncate toward zero; out-of-range values saturate like the Cortex-M4F VCVT
    (x86 would give INT_MIN - undefined behaviour in C either way, so designs that overflow Q2.30 are unpinned)."""
    return max(-2 ** 31, min(2 ** 31 - 1, int(math.trunc(v))))


def biquad_double_to_int(coef):
    """setCoefficients(stage, const double*): coef * 2^30, truncated (filter_biquad.h:44-52)."""
    return np.array([_c_int(float(c) * 1073741824.0) for c in coef], np.int32)


def _w0(frequency, fs):
    return float(f32(frequency)) * (2 * 3.141592654 / fs)


def biquad_lowpass(frequency, q=0.7071, fs=AUDIO_SAMPLE_RATE_EXACT):
    w0 = _w0(frequency, fs)
    sinW0, cosW0 = math.sin(w0), math.cos(w0)
    alpha = sinW0 / (float(f32(q)) * 2.0)
    scale = 1073741824.0 / (1.0 + alpha)
    b0 = _c_int(((1.0 - cosW0) / 2.0) * scale)
    return np.array([b0, _c_int((1.0 - cosW0) * scale), b0, _c_int((-2.0 * cosW0) * scale), _c_int((1.0 - alpha) * scale)], np.int32)


def biquad_highpass(frequency, q=0.7071, fs=AUDIO_SAMPLE_RATE_EXACT):
    w0 = _w0(frequency, fs)
    sinW0, cosW0 = math.sin(w0), math.cos(w0)
    alpha = sinW0 / (float(f32(q)) * 2.0)
    scale = 1073741824.0 / (1.0 + alpha)
    b0 = _c_int(((1.0 + cosW0) / 2.0) * scale)
    return np.array([b0, _c_int(-(1.0 + cosW0) * scale), b0, _c_int((-2.0 * cosW0) * scale), _c_int((1.0 - alpha) * scale)], np.int32)


def biquad_bandpass(frequency, q=1.0, fs=AUDIO_SAMPLE_RATE_EXACT):
    w0 = _w0(frequency, fs)
    sinW0, cosW0 = math.sin(w0), math.cos(w0)
    alpha = sinW0 / (float(f32(q)) * 2.0)
    scale = 1073741824.0 / (1.0 + alpha)
    return np.array([_c_int(alpha * scale), 0, _c_int((-alpha) * scale), _c_int((-2.0 * cosW0) * scale), _c_int((1.0 - alpha) * scale)], np.int32)


def biquad_notch(frequency, q=1.0, fs=AUDIO_SAMPLE_RATE_EXACT):
    w0 = _w0(frequency, fs)
    sinW0, cosW0 = math.sin(w0), math.cos(w0)
    alpha = sinW0 / (float(f32(q)) * 2.0)
    scale = 1073741824.0 / (1.0 + alpha)
    b0 = _c_int(scale)
    return np.array([b0, _c_int((-2.0 * cosW0) * scale), b0, _c_int((-2.0 * cosW0) * scale), _c_int((1.0 - alpha) * scale)], np.int32)


def _shelf_terms(frequency, gain, slope, fs):
    a = math.pow(10.0, float(f32(gain)) / 40.0)
    w0 = _w0(frequency, fs)
    sinW0, cosW0 = math.sin(w0), math.cos(w0)
    sinsq = sinW0 * math.sqrt((math.pow(a, 2.0) + 1.0) * (1.0 / float(f32(slope)) - 1.0) + 2.0 * a)
    return a, sinsq, (a - 1.0) * cosW0, (a + 1.0) * cosW0


def biquad_lowshelf(frequency, gain, slope=1.0, fs=AUDIO_SAMPLE_RATE_EXACT):
    a, sinsq, aMinus, aPlus = _shelf_terms(frequency, gain, slope, fs)
    scale = 1073741824.0 / ((a + 1.0) + aMinus + sinsq)
    return np.array([_c_int(a * ((a + 1.0) - aMinus + sinsq) * scale), _c_int(2.0 * a * ((a - 1.0) - aPlus) * scale),
                     _c_int(a * ((a + 1.0) - aMinus - sinsq) * scale), _c_int(-2.0 * ((a - 1.0) + aPlus) * scale),
                     _c_int(((a + 1.0) + aMinus - sinsq) * scale)], np.int32)


def biquad_highshelf(frequency, gain, slope=1.0, fs=AUDIO_SAMPLE_RATE_EXACT):
    a, sinsq, aMinus, aPlus = _shelf_terms(frequency, gain, slope, fs)
    scale = 1073741824.0 / ((a + 1.0) - aMinus + sinsq)
    return np.array([_c_int(a * ((a + 1.0) + aMinus + sinsq) * scale), _c_int(-2.0 * a * ((a - 1.0) + aPlus) * scale),
                     _c_int(a * ((a + 1.0) + aMinus - sinsq) * scale), _c_int(2.0 * ((a - 1.0) - aPlus) * scale),
                     _c_int(((a + 1.0) - aMinus - sinsq) * scale)], np.int32)
