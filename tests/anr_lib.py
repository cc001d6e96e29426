"""ctypes bindings for the ANR checkers (TEST INFRASTRUCTURE ONLY): the oracle's restatement (orc_anr_*, oracle/msdr_oracle.c) and
the reference's own block compiled from /root/reference (ref_anr_block, oracle/ref_anr.cpp)."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

import oracle_lib as ol

BLOCK = 128
_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


class OrcAnr:
    def __init__(self, n_channels):
        ol._ensure(ol.ORACLE_SO, "libmsdr_oracle.so")
        L = self.L = C.CDLL(ol.ORACLE_SO)
        L.orc_anr_new.restype = C.c_void_p
        L.orc_anr_new.argtypes = [C.c_uint32]
        L.orc_anr_free.argtypes = [C.c_void_p]
        L.orc_anr_run.argtypes = [C.c_void_p, C.c_uint32, C.c_int, _i16p, C.c_uint32, C.c_size_t]
        L.orc_anr_get.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int), _f32p, _f32p]
        self.n = n_channels
        self.h = L.orc_anr_new(n_channels)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_anr_free(self.h)
            self.h = None

    def run(self, mode, data):
        d = np.ascontiguousarray(data, np.int16).copy()
        self.L.orc_anr_run(self.h, self.n, int(mode), d, d.shape[1] // BLOCK, d.shape[1])
        return d

    def state(self, ch):
        lidx, ng, idx = C.c_float(), C.c_float(), C.c_int()
        w, d = np.zeros(64, np.float32), np.zeros(512, np.float32)
        self.L.orc_anr_get(self.h, ch, C.byref(lidx), C.byref(ng), C.byref(idx), w, d)
        return dict(lidx=np.float32(lidx.value), ngamma=np.float32(ng.value), in_idx=idx.value, w=w, d=d)


def ref_anr_run(mode, stream):
    """One channel through the reference's own block, in a fresh process (its state is in function statics)."""
    stream = np.ascontiguousarray(stream, np.int16)
    assert stream.ndim == 1 and stream.size % BLOCK == 0
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, "in.npy"), os.path.join(td, "out.npy")
        np.save(fi, stream)
        code = ("import ctypes as C, numpy as np\n"
                f"L = C.CDLL({ol.REF_SO!r})\n"
                "p = np.ctypeslib.ndpointer(dtype=np.int16, flags='C_CONTIGUOUS')\n"
                "L.ref_anr_block.argtypes = [C.c_int, p]\n"
                f"x = np.load({fi!r}).copy()\n"
                "for i in range(0, x.size, 128):\n"
                "    b = np.ascontiguousarray(x[i:i + 128]); L.ref_anr_block(" + str(int(mode)) + ", b); x[i:i + 128] = b\n"
                f"np.save({fo!r}, x)\n")
        subprocess.run([sys.executable, "-c", code], check=True)
        return np.load(fo)


def audio_stream(n_channels, n_samples, seed=0):
    """Demodulated-audio-like int16: speech-band noise plus steady tones ("birdies") the notch should remove."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples)
    out = np.empty((n_channels, n_samples), np.int16)
    for c in range(n_channels):
        noise = np.convolve(rng.normal(0, 1, n_samples + 15), np.ones(16) / 4.0, mode="valid") * (300 + 250 * (c % 7))
        tone = (2000 + 900 * (c % 5)) * np.sin(2 * np.pi * (0.031 + 0.007 * (c % 11)) * t + c)
        tone2 = (600 * ((c // 3) % 3)) * np.sin(2 * np.pi * 0.11 * t)
        out[c] = np.clip(np.round(noise + tone + tone2), -32768, 32767).astype(np.int16)
    return out
