"""Front-end conditioning object: the ADC DC-blocking high-pass, the `amp_adc` AudioAmplifier and the
sketch's AGC() in front of the receive chain, batched over channels.  Mirrors `AudioInputAnalog` (input_adc.cpp:198-212),
`AudioAmplifier` (mixer.{h,cpp}) and `AGC()` (Minimal-SDR.ino:445-515); all computation is in csrc/msdr_frontend.cu."""
import ctypes as C

import numpy as np

from . import capi

AGC_START, AGC_MAX = 0.25, 40.0  # Minimal-SDR.ino:94-95


class Frontend:
    def __init__(self, n_channels, device=0, agc_start=AGC_START, agc_max=AGC_MAX, agc_on=True):
        self._L = capi.lib()
        self.n_channels = int(n_channels)
        h = C.c_void_p()
        st = self._L.msdr_frontend_create(C.byref(h), device, self.n_channels, agc_start, agc_max, int(bool(agc_on)))
        if st != capi.OK:
            msg = self._L.msdr_frontend_last_error(None)
            raise capi.MsdrError(st, msg.decode() if msg else "")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self._L.msdr_frontend_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != capi.OK:
            msg = self._L.msdr_frontend_last_error(self.h)
            raise capi.MsdrError(st, msg.decode() if msg else "")
        return st

    def preset(self, first_reading, ch0=0, nch=None):
        """AudioInputAnalog::init: hpf_x1 = first ADC reading << 14, hpf_y1 = 0 (input_adc.cpp:59-63)."""
        nch = self.n_channels - ch0 if nch is None else nch
        return self._ck(self._L.msdr_frontend_preset(self.h, ch0, nch, int(first_reading)))

    def update(self, adc, out=None):
        """adc: uint16 [n_channels, n_blocks*128] raw codes (host) -> int16 conditioned IF samples."""
        adc = np.ascontiguousarray(adc, np.uint16)
        assert adc.ndim == 2 and adc.shape[0] == self.n_channels and adc.shape[1] % capi.BLOCK == 0
        if out is None:
            out = np.empty(adc.shape, np.int16)
        assert out.dtype == np.int16 and out.shape == adc.shape and out.flags["C_CONTIGUOUS"]
        self._ck(self._L.msdr_frontend_update(self.h, capi.ptr(adc), capi.ptr(out), adc.shape[1] // capi.BLOCK, adc.shape[1]))
        return out

    def update_device(self, d_adc, d_out, n_blocks, stride):
        return self._ck(self._L.msdr_frontend_update_device(self.h, C.c_void_p(int(d_adc)), C.c_void_p(int(d_out)), int(n_blocks), int(stride)))

    def set_stream(self, cuda_stream):
        # 0 = torch's default stream -> cudaStreamLegacy (handle 1); None = the object's own stream (see ReceiveChain.set_stream)
        h = 0 if cuda_stream is None else (int(cuda_stream) or 1)
        return self._ck(self._L.msdr_frontend_set_stream(self.h, C.c_void_p(h)))

    def set_option(self, key, value):
        return self._ck(self._L.msdr_frontend_set_option(self.h, key.encode(), int(value)))

    def synchronize(self):
        return self._ck(self._L.msdr_frontend_synchronize(self.h))

    def get_state(self, ch):
        st = capi.FrontendState()
        self._ck(self._L.msdr_frontend_get_state(self.h, int(ch), C.byref(st)))
        return st

    def set_state(self, ch, st):
        return self._ck(self._L.msdr_frontend_set_state(self.h, int(ch), C.byref(st)))

    def launch_count(self):
        return int(self._L.msdr_frontend_launch_count(self.h))


def amp_gain_multiplier(gain):
    return int(capi.lib().msdr_amp_gain_multiplier(float(gain)))
