"""The BASELINE.json configurations as data: channel counts, rates, tap tables, mode layout, biquad cascade, update sizes.

bench.py, the parity tests and the C5 scaling runs all take their shapes from here, so "C4" means one thing everywhere.
Host-side description only (tables come from the sketch's constants or the host-side designer); nothing is computed on
the data path.

  C3  batched 4096 channels x 10 s, mixed AM/SSB/CW, the sketch's tables (AM 102 taps, SSB/CW 86), 44.1 kHz  -- per GPU
  C4  wideband long-tap variant: 255-tap FIR pair (+ the zero arm_fir_init_q15 asks for, arm_fir_init_q15.c:55-64 -> 256),
      16 384 channels, 192 kHz, 1 s
  C5  2^20 channels x 60 s streaming, sharded over the GPUs of the job (strong scaling); the stream is processed in updates of
      k*128 samples with state carried; a bench step is a bounded slice of the 60 s (bench.py says how much)
"""
from dataclasses import dataclass, field

import numpy as np

from . import capi, design

BLOCK = capi.BLOCK


@dataclass
class Workload:
    name: str
    describe: str
    channels: int            # per GPU for weak scaling, total for strong scaling
    scaling: str             # "weak" | "strong"
    fs: float
    seconds: float           # signal length of one bench step, per channel
    blocks_per_update: int
    max_taps: int
    tables: dict = field(default_factory=dict)   # mode -> (cI, cQ) int16 arrays
    biquad1: np.ndarray = None
    biquad2: np.ndarray = None
    stream_seconds: float = None                  # C5: the whole stream the step is a slice of

    @property
    def blocks(self):
        n = int(round(self.seconds * self.fs))
        return (n + BLOCK - 1) // BLOCK

    def modes(self, n_channels, ch0=0):
        """mode = {AM, USB, LSB, CW}[c mod 4]"""
        tab = (capi.MODE_AM, capi.MODE_USB, capi.MODE_LSB, capi.MODE_CW)
        return [tab[(ch0 + c) % 4] for c in range(n_channels)]

    def tables_for(self, mode):
        return self.tables[capi.MODE_AM if mode == capi.MODE_SYNCAM else mode]

    def configure(self, chain, ch0=0):
        """Bind modes, tables and the live biquad cascade on a ReceiveChain (or anything with its setters) whose channel 0 is
        global channel ch0.  Channels of one mode form 4 strided families, set with one ranged call per channel run."""
        C = chain.n_channels
        modes = self.modes(C, ch0)
        chain.biquad_set_coefficients(0, 0, self.biquad1)
        chain.biquad_set_coefficients(1, 0, self.biquad2)
        # everything AM first (one ranged call), then one list call per other mode
        chain.set_mode(capi.MODE_AM)
        chain.fir_init(*self.tables_for(capi.MODE_AM))
        marr = np.asarray(modes)
        for md in sorted(set(modes) - {capi.MODE_AM}):
            ch = np.nonzero(marr == md)[0].astype(np.uint32)
            chain.set_mode_list(md, ch)
            chain.fir_init_list(*self.tables_for(md), ch)
        return modes


def _sketch_tables(K):
    am = np.array(K["FIR_AM_coeffs_bw2800_fs24000"], np.int16)
    ssb = (np.array(K["FIR_SSB_I_coeffs"], np.int16), np.array(K["FIR_SSB_Q_coeffs"], np.int16))
    cw = (np.array(K["FIR_CW_I_coeffs"], np.int16), np.array(K["FIR_CW_Q_coeffs"], np.int16))
    return {capi.MODE_AM: (am, am), capi.MODE_USB: ssb, capi.MODE_LSB: ssb, capi.MODE_CW: cw}


def _pad256(c255):
    """255 designed taps + the zero arm_fir_init_q15 requires for an odd count (arm_fir_init_q15.c:55-64)"""
    return np.concatenate([np.asarray(c255, np.int16), np.zeros(1, np.int16)])


_c4_cache = {}


def c4_tables(fs=192000.0):
    """255-tap Kaiser designs from the sketch's own designer (calc_FIR_coeffs, Minimal-SDR.ino:782-899) at 192 kHz: AM low-pass on
    both branches; SSB / CW band-pass for I with the time-reversed table for Q, the relation the sketch's +-45 degree tables
    have (Minimal-SDR.ino:119-128: Q = reverse(I))."""
    if fs not in _c4_cache:
        am = _pad256(design.calc_FIR_coeffs(255, 9000.0, 70.0, 0, 0.0, fs))
        ssb = design.calc_FIR_coeffs(255, 6000.0, 70.0, 2, 4000.0, fs)
        cw = design.calc_FIR_coeffs(255, 3000.0, 70.0, 2, 1000.0, fs)
        _c4_cache[fs] = {capi.MODE_AM: (am, am), capi.MODE_USB: (_pad256(ssb), _pad256(ssb[::-1])), capi.MODE_LSB: (_pad256(ssb), _pad256(ssb[::-1])),
                         capi.MODE_CW: (_pad256(cw), _pad256(cw[::-1]))}
    return _c4_cache[fs]


def get(name, K):
    """K = load_ref_constants()"""
    lp, notch = np.array(K["biquad1_lowpass_coef"], np.int32), np.array(K["biquad2_notch_coef"], np.int32)
    name = name.lower()
    if name == "c3":
        return Workload("c3", "C3: batched 4096 channels x 10 s mixed AM/SSB/CW per GPU, state carried across 128-sample blocks",
                        4096, "weak", 44100.0, 10.0, 1024, 102, _sketch_tables(K), lp, notch)
    if name == "c4":
        return Workload("c4", "C4: wideband long-tap variant, 255-tap FIR pair (+ zero pad = 256), 16 384 channels, fs = 192 kHz, 1 s, fused chain",
                        16384, "weak", 192000.0, 1.0, 375, 256, c4_tables(), lp, notch)
    if name == "c5":
        return Workload("c5", "C5: 2^20 channels streaming, sharded over the GPUs of the job; a step is a slice of the 60 s stream",
                        1 << 20, "strong", 44100.0, 128 * BLOCK / 44100.0, 32, 102, _sketch_tables(K), lp, notch, stream_seconds=60.0)
    raise ValueError(f"unknown workload {name!r} (c3, c4, c5)")


def sample_channels(n_channels, want=48, seed=5):
    """Channels a parity check looks at: both ends, the edges of the 32-channel groups and 128-row tiles, and a seeded random rest."""
    edges = [0, 1, 2, 3, 30, 31, 32, 33, 63, 64, 126, 127, 128, 129, 255, 256, n_channels // 2 - 1, n_channels // 2, n_channels // 2 + 31,
             n_channels - 130, n_channels - 129, n_channels - 128, n_channels - 34, n_channels - 33, n_channels - 32, n_channels - 2, n_channels - 1]
    s = sorted({c for c in edges if 0 <= c < n_channels})
    rng = np.random.default_rng(seed)
    while len(s) < min(want, n_channels):
        c = int(rng.integers(0, n_channels))
        if c not in s:
            s.append(c)
    return sorted(s)
