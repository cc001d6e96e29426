"""ctypes bindings for the CPU checkers (TEST INFRASTRUCTURE ONLY).

Two libraries share one API shape, differing by symbol prefix:
  * ``orc_`` — oracle/libmsdr_oracle.so, the plain-C restatement (oracle/msdr_oracle.c); always available.
  * ``ref_`` — oracle/_ref/libmsdr_ref.so, the reference's own sources compiled with shims
               (oracle/Makefile); available where /root/reference was present at build time
               (this container) or where the prebuilt .so travelled (the GPU box).
Nothing under minimal-sdr_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libmsdr_oracle.so")
REF_SO = os.environ.get("MSDR_REF_SO") or os.path.join(ORACLE_DIR, "_ref", "libmsdr_ref.so")  # env: test another build of the same sources
# builds of the reference sources at other optimisation levels (oracle/Makefile `ref_variants`): CPU-baseline timing and a
# cross-check that the compiled reference does not depend on the optimiser
REF_VARIANTS = {"O2": REF_SO,
                "O2-novec": os.path.join(ORACLE_DIR, "_ref", "o2novec", "libmsdr_ref.so"),
                "O3-x86-64-v3": os.path.join(ORACLE_DIR, "_ref", "o3", "libmsdr_ref.so")}

MODE_SYNCAM, MODE_AM, MODE_LSB, MODE_USB, MODE_CW = 0, 1, 2, 3, 4  # stations.h:4
DEMOD_LSB, DEMOD_USB, DEMOD_AM_F32, DEMOD_AM_Q31 = 0, 1, 2, 3

_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(target="all"):
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, target], check=True)


def _ensure(path, target):
    if not os.path.exists(path):
        build(target)
    return os.path.exists(path)


def have_ref():
    if os.path.exists(REF_SO):
        return True
    if os.path.exists("/root/reference/Minimal-SDR.ino"):
        build("ref")
    return os.path.exists(REF_SO)


class CheckerLib:
    """Uniform view of either checker library."""

    def __init__(self, prefix, path=None):
        assert prefix in ("orc", "ref")
        self.prefix = prefix
        if prefix == "orc":
            _ensure(ORACLE_SO, "libmsdr_oracle.so")
            self.lib = C.CDLL(ORACLE_SO)
        elif path:  # another build of the same reference sources (REF_VARIANTS)
            self.lib = C.CDLL(path)
        else:
            if not have_ref():
                raise FileNotFoundError(REF_SO)
            self.lib = C.CDLL(REF_SO)
        L, p = self.lib, prefix + "_"
        self._f = {}

        def bind(name, restype, argtypes):
            fn = getattr(L, p + name)
            fn.restype, fn.argtypes = restype, argtypes
            self._f[name] = fn

        bind("mix_fs4", None, [_i16p, _i16p, _i16p, C.c_uint32])
        bind("fir_new", C.c_void_p, [C.c_uint16, _i16p, C.c_uint32, C.POINTER(C.c_int)])
        bind("fir_set_coefficients", None, [C.c_void_p, _i16p])
        bind("fir_free", None, [C.c_void_p])
        bind("fir_run", None, [C.c_void_p, _i16p, _i16p, C.c_uint32])
        bind("fir_state", C.POINTER(C.c_int16), [C.c_void_p])
        bind("sqrt_q31", C.c_int, [C.c_int32, C.POINTER(C.c_int32)])
        bind("demod", None, [C.c_int, _i16p, _i16p, _i16p, C.c_uint32])
        bind("biquad_new", C.c_void_p, [])
        bind("biquad_free", None, [C.c_void_p])
        bind("biquad_set_coefficients", None, [C.c_void_p, C.c_uint32, _i32p])
        bind("biquad_get_definition", None, [C.c_void_p, _i32p])
        bind("biquad_set_definition", None, [C.c_void_p, _i32p])
        if prefix == "orc":
            bind("biquad_update", None, [C.c_void_p, _i16p, C.c_uint32])
            bind("freq_conv", None, [C.c_int, C.c_int, _i16p, _i16p, _i16p, _i16p, C.c_uint32])
        else:
            bind("biquad_update", None, [C.c_void_p, _i16p])
            bind("biquad_design", None, [C.c_void_p, C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_float])
            bind("calc_FIR_coeffs", None, [_i16p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float])
            bind("audio_sample_rate_exact", C.c_double, [])
            try:  # A6: the reference's own freq_conv.cpp + the vendored saturating primitives (oracle/ref_freq_conv.cpp, ref_q15_prims.c)
                bind("freq_conv", None, [C.c_int, C.c_int, _i16p, _i16p, _i16p, _i16p, C.c_uint32])
                bind("freq_conv_ex", C.c_int, [C.c_int, C.c_int, _i16p, _i16p, _i16p, _i16p, C.c_uint32, C.c_int, C.c_int, C.c_int])
                bind("qadd16", C.c_uint32, [C.c_uint32, C.c_uint32])
                bind("qsub16", C.c_uint32, [C.c_uint32, C.c_uint32])
                bind("clip_q31_to_q15", C.c_int16, [C.c_int32])
            except AttributeError:  # a _ref library built before these were added
                pass
        bind("chain_new", C.c_void_p, [C.c_uint32, C.c_int])
        bind("chain_free", None, [C.c_void_p])
        bind("chain_set_mode", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int])
        bind("chain_fir_init", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint16, _i16p, _i16p])
        bind("chain_fir_set_coefficients", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, _i16p, _i16p])
        bind("chain_biquad_set_coefficients", C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, _i32p])
        bind("chain_run", C.c_int, [C.c_void_p, _i16p, _i16p, C.c_uint32, C.c_size_t, C.c_int])

    # ---- stage level -------------------------------------------------------------------------
    def mix_fs4(self, x):
        x = np.ascontiguousarray(x, np.int16)
        I, Q = np.empty_like(x), np.empty_like(x)
        self._f["mix_fs4"](x, I, Q, x.size)
        return I, Q

    def fir(self, coeffs, x, block=128, hist_out=False, recoef=None):
        """Run one FIR over x in calls of `block` samples. recoef = (sample_index, new_coeffs) swaps taps in place."""
        coeffs = np.ascontiguousarray(coeffs, np.int16)
        x = np.ascontiguousarray(x, np.int16)
        st = C.c_int(0)
        h = self._f["fir_new"](len(coeffs), coeffs, block, C.byref(st))
        y = np.empty_like(x)
        if recoef is None:
            self._f["fir_run"](h, x, y, x.size)
        else:
            at, c2 = recoef
            ya, yb = np.empty(at, np.int16), np.empty(x.size - at, np.int16)
            self._f["fir_run"](h, np.ascontiguousarray(x[:at]), ya, at)
            self._f["fir_set_coefficients"](h, np.ascontiguousarray(c2, np.int16))
            self._f["fir_run"](h, np.ascontiguousarray(x[at:]), yb, x.size - at)
            y = np.concatenate([ya, yb])
        hist = np.ctypeslib.as_array(self._f["fir_state"](h), shape=(len(coeffs) - 1,)).copy()
        self._f["fir_free"](h)
        return (y, hist, st.value) if hist_out else y

    def fir_init_status(self, ntaps):
        st = C.c_int(0)
        h = self._f["fir_new"](ntaps, np.zeros(ntaps + 2, np.int16), 128, C.byref(st))
        self._f["fir_free"](h)
        return st.value

    def sqrt_q31(self, v):
        out = C.c_int32(0)
        st = self._f["sqrt_q31"](int(v), C.byref(out))
        return out.value, st

    def demod(self, kind, I, Q):
        I = np.ascontiguousarray(I, np.int16)
        Q = np.ascontiguousarray(Q, np.int16)
        out = np.zeros_like(I)
        self._f["demod"](kind, I, Q, out, I.size)
        return out

    def biquad(self, stages, x, definition_out=False, definition_in=None):
        """stages: list of (stage_index, int32[5]) applied in order; x processed in 128-sample updates."""
        x = np.ascontiguousarray(x, np.int16).copy()
        assert x.size % 128 == 0
        h = self._f["biquad_new"]()
        if definition_in is not None:
            self._f["biquad_set_definition"](h, np.ascontiguousarray(definition_in, np.int32))
        for s, coef in stages:
            self._f["biquad_set_coefficients"](h, s, np.ascontiguousarray(coef, np.int32))
        for b in range(x.size // 128):
            blk = np.ascontiguousarray(x[b * 128:(b + 1) * 128])
            if self.prefix == "orc":
                self._f["biquad_update"](h, blk, 128)
            else:
                self._f["biquad_update"](h, blk)
            x[b * 128:(b + 1) * 128] = blk
        d = np.zeros(32, np.int32)
        self._f["biquad_get_definition"](h, d)
        self._f["biquad_free"](h)
        return (x, d) if definition_out else x

    def freq_conv(self, direction, passthrough, I, Q, oscI, oscQ):
        I = np.ascontiguousarray(I, np.int16).copy()
        Q = np.ascontiguousarray(Q, np.int16).copy()
        self._f["freq_conv"](int(direction), int(passthrough), I, Q,
                             np.ascontiguousarray(oscI, np.int16), np.ascontiguousarray(oscQ, np.int16), I.size)
        return I, Q

    def freq_conv_ex(self, direction, passthrough, I, Q, oscI, oscQ, have_I=1, have_Q=1, fail_alloc=0):
        """compiled reference only: (I, Q, transmitted) with a missing input block or a failed allocate()"""
        I = np.ascontiguousarray(I, np.int16).copy()
        Q = np.ascontiguousarray(Q, np.int16).copy()
        ok = self._f["freq_conv_ex"](int(direction), int(passthrough), I, Q, np.ascontiguousarray(oscI, np.int16),
                                     np.ascontiguousarray(oscQ, np.int16), I.size, have_I, have_Q, fail_alloc)
        return I, Q, bool(ok)

    def biquad_design(self, kind, frequency, p2, p3=1.0, stage=0):
        h = self._f["biquad_new"]()
        self._f["biquad_design"](h, kind, stage, frequency, p2, p3)
        d = np.zeros(32, np.int32)
        self._f["biquad_get_definition"](h, d)
        self._f["biquad_free"](h)
        c = d[stage * 8:stage * 8 + 5].copy()
        c[3], c[4] = -c[3], -c[4]  # stored negated (filter_biquad.cpp:93-94)
        return c

    def calc_fir_coeffs(self, n, fc, astop, ftype, dfc, fs):
        out = np.zeros(2 * n + 8, np.int16)
        self._f["calc_FIR_coeffs"](out, n, fc, astop, ftype, dfc, fs)
        return out[:n].copy()

    # ---- chain level -------------------------------------------------------------------------
    def chain(self, n_channels, am_q31=False):
        return CheckerChain(self, n_channels, am_q31)


class CheckerChain:
    def __init__(self, lib, n_channels, am_q31):
        self.L, self.n = lib, n_channels
        self.h = lib._f["chain_new"](n_channels, int(am_q31))

    def close(self):
        if self.h:
            self.L._f["chain_free"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mode(self, ch0, nch, mode):
        return self.L._f["chain_set_mode"](self.h, ch0, nch, mode)

    def fir_init(self, ch0, nch, cI, cQ):
        cI = np.ascontiguousarray(cI, np.int16)
        cQ = np.ascontiguousarray(cQ, np.int16)
        return self.L._f["chain_fir_init"](self.h, ch0, nch, len(cI), cI, cQ)

    def fir_set_coefficients(self, ch0, nch, cI, cQ):
        return self.L._f["chain_fir_set_coefficients"](self.h, ch0, nch, np.ascontiguousarray(cI, np.int16),
                                                       np.ascontiguousarray(cQ, np.int16))

    def set_anr(self, ch0, nch, anr_on):
        """ANR_on (Minimal-SDR.ino:99); only the plain-C oracle has it in its chain (the compiled reference's block is pinned separately)."""
        f = self.L.lib.orc_chain_set_anr
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        return f(self.h, ch0, nch, int(anr_on))

    def biquad_set_coefficients(self, obj, ch0, nch, stage, coef):
        return self.L._f["chain_biquad_set_coefficients"](self.h, obj, ch0, nch, stage,
                                                          np.ascontiguousarray(coef, np.int32))

    def run(self, x, n_threads=0):
        """x: int16 [n_channels, n_blocks*128] -> (out, threads_used)"""
        x = np.ascontiguousarray(x, np.int16)
        assert x.shape[0] == self.n and x.shape[1] % 128 == 0
        out = np.zeros_like(x)
        used = self.L._f["chain_run"](self.h, x, out, x.shape[1] // 128, x.shape[1], n_threads)
        return out, used
