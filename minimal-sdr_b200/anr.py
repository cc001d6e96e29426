"""LMS automatic notch / noise reduction object: the `ANR_on > 0` block of the sketch's demodulation()
(Minimal-SDR.ino:702-770), batched over channels.  All computation is in csrc/msdr_anr.cu."""
import ctypes as C

import numpy as np

from . import capi

NOTCH, NOISE_REDUCTION = 1, 2  # ANR_on values, Minimal-SDR.ino:99


class Anr:
    def __init__(self, n_channels, device=0):
        self._L = capi.lib()
        self.n_channels = int(n_channels)
        h = C.c_void_p()
        st = self._L.msdr_anr_create(C.byref(h), device, self.n_channels)
        if st != capi.OK:
            msg = self._L.msdr_anr_last_error(None)
            raise capi.MsdrError(st, msg.decode() if msg else "")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self._L.msdr_anr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != capi.OK:
            msg = self._L.msdr_anr_last_error(self.h)
            raise capi.MsdrError(st, msg.decode() if msg else "")
        return st

    def update(self, mode, audio):
        """audio: int16 [n_channels, n_blocks*128] demodulated audio (host) -> filtered copy."""
        out = np.ascontiguousarray(audio, np.int16).copy()
        assert out.ndim == 2 and out.shape[0] == self.n_channels and out.shape[1] % capi.BLOCK == 0
        self._ck(self._L.msdr_anr_update(self.h, int(mode), capi.ptr(out), out.shape[1] // capi.BLOCK, out.shape[1]))
        return out

    def update_device(self, mode, d_data, n_blocks, stride):
        return self._ck(self._L.msdr_anr_update_device(self.h, int(mode), C.c_void_p(int(d_data)), int(n_blocks), int(stride)))

    def set_stream(self, cuda_stream):
        # 0 = torch's default stream -> cudaStreamLegacy (handle 1); None = the object's own stream (see ReceiveChain.set_stream)
        h = 0 if cuda_stream is None else (int(cuda_stream) or 1)
        return self._ck(self._L.msdr_anr_set_stream(self.h, C.c_void_p(h)))

    def synchronize(self):
        return self._ck(self._L.msdr_anr_synchronize(self.h))

    def get_state(self, ch):
        st = capi.AnrState()
        self._ck(self._L.msdr_anr_get_state(self.h, int(ch), C.byref(st)))
        return st

    def set_state(self, ch, st):
        return self._ck(self._L.msdr_anr_set_state(self.h, int(ch), C.byref(st)))

    def launch_count(self):
        return int(self._L.msdr_anr_launch_count(self.h))
