"""Synchronous-AM demodulator with PLL: `case SYNCAM` of the sketch's demodulation switch
(Minimal-SDR.ino:631-688), batched over channels.  All computation is in csrc/msdr_syncam.cu."""
import ctypes as C

import numpy as np

from . import capi


class SyncAm:
    def __init__(self, n_channels, device=0):
        self._L = capi.lib()
        self.n_channels = int(n_channels)
        h = C.c_void_p()
        st = self._L.msdr_syncam_create(C.byref(h), device, self.n_channels)
        if st != capi.OK:
            msg = self._L.msdr_syncam_last_error(None)
            raise capi.MsdrError(st, msg.decode() if msg else "")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self._L.msdr_syncam_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != capi.OK:
            msg = self._L.msdr_syncam_last_error(self.h)
            raise capi.MsdrError(st, msg.decode() if msg else "")
        return st

    def update(self, I, Q):
        """I, Q: int16 [n_channels, n_blocks*128] FIR-filtered baseband (host) -> int16 audio."""
        I = np.ascontiguousarray(I, np.int16)
        Q = np.ascontiguousarray(Q, np.int16)
        assert I.shape == Q.shape and I.ndim == 2 and I.shape[0] == self.n_channels and I.shape[1] % capi.BLOCK == 0
        out = np.empty(I.shape, np.int16)
        self._ck(self._L.msdr_syncam_update(self.h, capi.ptr(I), capi.ptr(Q), capi.ptr(out), I.shape[1] // capi.BLOCK, I.shape[1]))
        return out

    def get_state(self, ch):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._ck(self._L.msdr_syncam_get_state(self.h, int(ch), C.byref(a), C.byref(b), C.byref(c)))
        return np.float32(a.value), np.float32(b.value), np.float32(c.value)

    def set_state(self, ch, fil_out, omega2, phzerror):
        return self._ck(self._L.msdr_syncam_set_state(self.h, int(ch), float(fil_out), float(omega2), float(phzerror)))


def constants():
    """(omega_min, omega_max, g1, g2) as float32, evaluated by the library with the sketch's own expressions (host code, no GPU)."""
    v = [C.c_float() for _ in range(4)]
    capi.lib().msdr_syncam_constants(*[C.byref(x) for x in v])
    return tuple(np.float32(x.value) for x in v)
