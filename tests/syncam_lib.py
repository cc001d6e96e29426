"""ctypes bindings for the synchronous-AM checkers (TEST INFRASTRUCTURE ONLY): the oracle's restatement (orc_syncam_*) and the
reference's own `case SYNCAM` arm compiled from /root/reference (ref_syncam_block, oracle/ref_syncam.cpp)."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

import oracle_lib as ol

BLOCK = 128
_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")


class OrcSyncAm:
    def __init__(self, n_channels):
        ol._ensure(ol.ORACLE_SO, "libmsdr_oracle.so")
        L = self.L = C.CDLL(ol.ORACLE_SO)
        L.orc_syncam_new.restype = C.c_void_p
        L.orc_syncam_new.argtypes = [C.c_uint32]
        L.orc_syncam_free.argtypes = [C.c_void_p]
        L.orc_syncam_run.argtypes = [C.c_void_p, C.c_uint32, _i16p, _i16p, _i16p, C.c_uint32, C.c_size_t]
        L.orc_syncam_get.argtypes = [C.c_void_p, C.c_uint32] + [C.POINTER(C.c_float)] * 3
        L.orc_syncam_constants.argtypes = [C.POINTER(C.c_float)] * 4
        self.n = n_channels
        self.h = L.orc_syncam_new(n_channels)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_syncam_free(self.h)
            self.h = None

    def run(self, I, Q):
        I, Q = np.ascontiguousarray(I, np.int16), np.ascontiguousarray(Q, np.int16)
        out = np.empty(I.shape, np.int16)
        self.L.orc_syncam_run(self.h, self.n, I, Q, out, I.shape[1] // BLOCK, I.shape[1])
        return out

    def state(self, ch):
        v = [C.c_float() for _ in range(3)]
        self.L.orc_syncam_get(self.h, ch, *[C.byref(x) for x in v])
        return tuple(np.float32(x.value) for x in v)

    def constants(self):
        v = [C.c_float() for _ in range(4)]
        self.L.orc_syncam_constants(*[C.byref(x) for x in v])
        return tuple(np.float32(x.value) for x in v)


def ref_syncam_run(I, Q):
    """One channel through the reference's own arm, in a fresh process (its state is in function statics)."""
    I, Q = np.ascontiguousarray(I, np.int16), np.ascontiguousarray(Q, np.int16)
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "i.npy"), I)
        np.save(os.path.join(td, "q.npy"), Q)
        code = ("import ctypes as C, numpy as np\n"
                f"L = C.CDLL({ol.REF_SO!r}); p = np.ctypeslib.ndpointer(dtype=np.int16, flags='C_CONTIGUOUS')\n"
                "L.ref_syncam_block.argtypes = [p, p, p]\n"
                f"I = np.load({td!r} + '/i.npy'); Q = np.load({td!r} + '/q.npy'); o = np.zeros_like(I)\n"
                "for k in range(0, I.size, 128):\n"
                "    a = np.ascontiguousarray(I[k:k + 128]); b = np.ascontiguousarray(Q[k:k + 128]); c = np.zeros(128, np.int16)\n"
                "    L.ref_syncam_block(a, b, c); o[k:k + 128] = c\n"
                f"np.save({td!r} + '/o.npy', o)\n")
        subprocess.run([sys.executable, "-c", code], check=True)
        return np.load(os.path.join(td, "o.npy"))


def baseband(n_channels, n_samples, seed=0):
    """AM carriers a few hundred Hz (at 24 kHz) off zero with tone + noise modulation: filtered I/Q as the FIR pair delivers it."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples)
    I = np.empty((n_channels, n_samples), np.int16)
    Q = np.empty((n_channels, n_samples), np.int16)
    for c in range(n_channels):
        off = (-350 + 90 * (c % 9)) / 24000.0
        env = 1 + 0.5 * np.sin(2 * np.pi * t * (0.004 + 0.001 * (c % 4))) + 0.2 * np.sin(2 * np.pi * t * 0.021)
        amp = 1500 + 900 * (c % 7)
        ph = 2 * np.pi * off * t + 0.3 * c
        I[c] = np.round(amp * env * np.cos(ph) + rng.normal(0, 25, n_samples)).clip(-32768, 32767).astype(np.int16)
        Q[c] = np.round(amp * env * np.sin(ph) + rng.normal(0, 25, n_samples)).clip(-32768, 32767).astype(np.int16)
    return I, Q
