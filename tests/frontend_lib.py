"""ctypes bindings for the front-end checkers (TEST INFRASTRUCTURE ONLY): the oracle's restatement (orc_*, oracle/msdr_oracle.c)
and the reference's own code compiled from /root/reference (ref_*, oracle/ref_frontend.cpp + oracle/Makefile)."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

import oracle_lib as ol

BLOCK = 128
_u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_i32 = C.POINTER(C.c_int32)


class OrcAgc(C.Structure):
    _fields_ = [("buf", C.c_int16 * 25), ("idx", C.c_int), ("val", C.c_float), ("max", C.c_float), ("on", C.c_int)]


class Orc:
    def __init__(self):
        ol._ensure(ol.ORACLE_SO, "libmsdr_oracle.so")
        L = self.L = C.CDLL(ol.ORACLE_SO)
        L.orc_adc_hpf.argtypes = [_u16p, _i16p, C.c_uint32, _i32, _i32]
        L.orc_amp_multiplier.restype = C.c_int32
        L.orc_amp_multiplier.argtypes = [C.c_float]
        L.orc_amp_apply.restype = C.c_int
        L.orc_amp_apply.argtypes = [_i16p, C.c_uint32, C.c_int32]
        L.orc_agc_init.argtypes = [C.POINTER(OrcAgc), C.c_float, C.c_float, C.c_int]
        L.orc_agc_absmax.restype = C.c_uint16
        L.orc_agc_absmax.argtypes = [_i16p]
        L.orc_agc_update.restype = C.c_int
        L.orc_agc_update.argtypes = [C.POINTER(OrcAgc), C.c_uint16, _i32]
        L.orc_frontend_new.restype = C.c_void_p
        L.orc_frontend_new.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_int]
        L.orc_frontend_free.argtypes = [C.c_void_p]
        L.orc_frontend_preset.argtypes = [C.c_void_p, C.c_uint32, C.c_uint16]
        L.orc_frontend_get.argtypes = [C.c_void_p, C.c_uint32, _i32, _i32, _i32, C.POINTER(C.c_float), C.POINTER(C.c_int), _i16p]
        L.orc_frontend_run.argtypes = [C.c_void_p, C.c_uint32, _u16p, _i16p, C.c_uint32, C.c_size_t]

    def hpf(self, codes, x1=0, y1=0):
        codes = np.ascontiguousarray(codes, np.uint16)
        out = np.empty(codes.shape, np.int16)
        a, b = C.c_int32(x1), C.c_int32(y1)
        self.L.orc_adc_hpf(codes, out, codes.size, C.byref(a), C.byref(b))
        return out, a.value, b.value

    def amp_multiplier(self, gain):
        return int(self.L.orc_amp_multiplier(float(gain)))

    def amp_apply(self, data, mult):
        d = np.ascontiguousarray(data, np.int16).copy()
        sent = self.L.orc_amp_apply(d, d.size, int(mult))
        return d, sent

    def agc_trajectory(self, blocks, start=0.25, mx=40.0, on=1):
        """blocks: int16 [n, 128] -> (agc_val float32 [n], multiplier int32 [n]) after each block."""
        a = OrcAgc()
        self.L.orc_agc_init(C.byref(a), start, mx, on)
        mult = self.amp_multiplier(start)
        vals, mults = [], []
        for blk in np.ascontiguousarray(blocks, np.int16):
            m = C.c_int32(mult)
            if self.L.orc_agc_update(C.byref(a), self.L.orc_agc_absmax(np.ascontiguousarray(blk)), C.byref(m)):
                mult = m.value
            vals.append(a.val)
            mults.append(mult)
        return np.array(vals, np.float32), np.array(mults, np.int32)

    def frontend(self, n_channels, start=0.25, mx=40.0, on=1):
        return OrcFrontend(self, n_channels, start, mx, on)


class OrcFrontend:
    def __init__(self, orc, n, start, mx, on):
        self.o, self.n = orc, n
        self.h = orc.L.orc_frontend_new(n, start, mx, int(on))

    def __del__(self):
        if getattr(self, "h", None):
            self.o.L.orc_frontend_free(self.h)
            self.h = None

    def preset(self, ch, first):
        self.o.L.orc_frontend_preset(self.h, ch, int(first))

    def run(self, codes):
        codes = np.ascontiguousarray(codes, np.uint16)
        out = np.empty(codes.shape, np.int16)
        self.o.L.orc_frontend_run(self.h, self.n, codes, out, codes.shape[1] // BLOCK, codes.shape[1])
        return out

    def state(self, ch):
        x1, y1, m, idx = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int()
        v = C.c_float()
        buf = np.zeros(25, np.int16)
        self.o.L.orc_frontend_get(self.h, ch, C.byref(x1), C.byref(y1), C.byref(m), C.byref(v), C.byref(idx), buf)
        return dict(hpf_x1=x1.value, hpf_y1=y1.value, multiplier=m.value, agc_val=np.float32(v.value), agc_idx=idx.value, agc_buffer=buf)


class Ref:
    """The reference's own front-end code (oracle/_ref/libmsdr_ref.so)."""

    def __init__(self):
        if not ol.have_ref():
            raise FileNotFoundError(ol.REF_SO)
        L = self.L = C.CDLL(ol.REF_SO)
        L.ref_adc_hpf_block.argtypes = [_i16p, _i32, _i32]
        L.ref_amp_multiplier.restype = C.c_int32
        L.ref_amp_multiplier.argtypes = [C.c_float]
        L.ref_amp_block.restype = C.c_int
        L.ref_amp_block.argtypes = [C.c_float, _i16p]

    def hpf(self, codes, x1=0, y1=0):
        codes = np.ascontiguousarray(codes, np.uint16)
        assert codes.size % BLOCK == 0
        data = codes.view(np.int16).copy()
        a, b = C.c_int32(x1), C.c_int32(y1)
        for i in range(0, data.size, BLOCK):
            blk = np.ascontiguousarray(data[i:i + BLOCK])
            self.L.ref_adc_hpf_block(blk, C.byref(a), C.byref(b))
            data[i:i + BLOCK] = blk
        return data, a.value, b.value

    def amp_multiplier(self, gain):
        return int(self.L.ref_amp_multiplier(float(gain)))

    def amp_block(self, gain, data):
        d = np.ascontiguousarray(data, np.int16).copy()
        sent = self.L.ref_amp_block(float(gain), d)
        return d, sent

    @staticmethod
    def agc_trajectory(blocks, start=0.25, mx=40.0, on=1):
        """AGC() keeps its history in function statics: every trajectory runs in a fresh process."""
        blocks = np.ascontiguousarray(blocks, np.int16)
        with tempfile.TemporaryDirectory() as td:
            fi, fo = os.path.join(td, "in.npy"), os.path.join(td, "out.npz")
            np.save(fi, blocks)
            code = (
                "import ctypes as C, numpy as np, sys\n"
                f"L = C.CDLL({ol.REF_SO!r})\n"
                "L.ref_agc_config.argtypes = [C.c_float, C.c_float, C.c_int]\n"
                "p = np.ctypeslib.ndpointer(dtype=np.int16, flags='C_CONTIGUOUS')\n"
                "L.ref_agc_block.argtypes = [p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]\n"
                f"b = np.load({fi!r}); L.ref_agc_config({start!r}, {mx!r}, {int(on)})\n"
                "v, m = C.c_float(), C.c_int32(); vs, ms = [], []\n"
                "for blk in b:\n"
                "    blk = np.ascontiguousarray(blk); L.ref_agc_block(blk, C.byref(v), C.byref(m)); vs.append(v.value); ms.append(m.value)\n"
                f"np.savez({fo!r}, v=np.array(vs, np.float32), m=np.array(ms, np.int32))\n")
            subprocess.run([sys.executable, "-c", code], check=True)
            z = np.load(fo)
            return z["v"], z["m"]


def adc_stream(n_channels, n_samples, seed=0, bits=12):
    """Synthetic raw ADC codes: DC offset + carrier at fs/4 with slowly varying, channel-dependent amplitude + noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples)
    out = np.empty((n_channels, n_samples), np.uint16)
    full = (1 << bits) - 1
    for c in range(n_channels):
        amp = (0.02 + 0.45 * ((c * 7) % 10) / 10.0) * (1.0 + 0.8 * np.sin(2 * np.pi * t / (9000.0 + 700 * c)))
        x = 0.5 + 0.03 * ((c % 5) - 2) + amp * np.sin(2 * np.pi * 0.25 * t + c) * 0.5 + rng.normal(0, 0.003, n_samples)
        out[c] = np.clip(np.round(x * full), 0, full).astype(np.uint16)
    return out
