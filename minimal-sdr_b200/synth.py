"""Deterministic synthetic int16 IF streams.  numpy on the host (tests, CPU baseline) and a torch
variant that generates the same kind of signal directly in HBM for the benchmark.

Per channel c (seed = 0x4D534452 ^ c):
  AM  : x[n] = round(A (1 + 0.5 sin 2 pi fm n/fs) cos(2 pi (fs/4 + d) n/fs) + w[n]),  A in [2000, 20000],
        fm in {400, 1000} Hz, d in [-200, 200] Hz, w uniform +-32 LSB
  SSB : three tones at fs/4 +- {300..2700} Hz (upper side for USB, lower for LSB), peak <= 24000
  CW  : one tone 800 Hz above fs/4, keyed on/off at 10 Hz
"""
import numpy as np

from . import capi

SEED = 0x4D534452


def channel_params(c):
    rng = np.random.default_rng(SEED ^ int(c))
    return {
        "A": rng.uniform(2000.0, 20000.0),
        "fm": (400.0, 1000.0)[int(rng.integers(0, 2))],
        "delta": rng.uniform(-200.0, 200.0),
        "tones": rng.uniform(300.0, 2700.0, 3),
        "phases": rng.uniform(0, 2 * np.pi, 3),
        "noise_seed": int(rng.integers(0, 2 ** 31)),
    }


def channel_stream(c, mode, n, fs=44100.0, n0=0):
    """int16[n] for channel c starting at absolute sample n0."""
    p = channel_params(c)
    t = (np.arange(n0, n0 + n, dtype=np.float64)) / fs
    w = np.random.default_rng(p["noise_seed"] + n0).uniform(-32.0, 32.0, n)
    f0 = fs / 4.0
    if mode in (capi.MODE_AM, capi.MODE_SYNCAM):
        x = p["A"] * (1.0 + 0.5 * np.sin(2 * np.pi * p["fm"] * t)) * np.cos(2 * np.pi * (f0 + p["delta"]) * t)
    elif mode in (capi.MODE_USB, capi.MODE_LSB):
        sgn = 1.0 if mode == capi.MODE_USB else -1.0
        x = sum(8000.0 * np.cos(2 * np.pi * (f0 + sgn * f) * t + ph) for f, ph in zip(p["tones"], p["phases"]))
    else:  # CW
        key = (np.floor(t * 10.0) % 2 == 0).astype(np.float64)
        x = p["A"] * key * np.cos(2 * np.pi * (f0 + 800.0) * t)
    return np.clip(np.rint(x + w), -32768, 32767).astype(np.int16)


def batch(modes, n, fs=44100.0, n0=0, ch0=0):
    """int16[len(modes), n]; row r is channel ch0 + r in mode modes[r]."""
    return np.stack([channel_stream(ch0 + r, m, n, fs, n0) for r, m in enumerate(modes)])


def mixed_modes(n_channels, ch0=0):
    """C3/C5 layout: mode = {AM, USB, LSB, CW}[c mod 4]."""
    tab = (capi.MODE_AM, capi.MODE_USB, capi.MODE_LSB, capi.MODE_CW)
    return [tab[(ch0 + c) % 4] for c in range(n_channels)]


def torch_batch(n_channels, n, device, fs=44100.0, ch0=0, n0=0):
    """Same families of signals generated in HBM with torch (fp32 phase, so not sample-identical to the numpy
    generator; the benchmark only needs realistic multi-tone int16 content).  Returns int16 [n_channels, n]."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(SEED ^ int(ch0))
    c = torch.arange(ch0, ch0 + n_channels, device=device)
    A = torch.empty(n_channels, 1, device=device).uniform_(2000.0, 20000.0, generator=g)
    fm = torch.where(torch.rand(n_channels, 1, device=device, generator=g) < 0.5, 400.0, 1000.0)
    delta = torch.empty(n_channels, 1, device=device).uniform_(-200.0, 200.0, generator=g)
    tones = torch.empty(n_channels, 3, device=device).uniform_(300.0, 2700.0, generator=g)
    t = (torch.arange(n0, n0 + n, device=device, dtype=torch.float64) / fs)
    tw = (2 * torch.pi * t)
    f0 = fs / 4.0
    mode = (c % 4).view(-1, 1)  # 0 AM, 1 USB, 2 LSB, 3 CW  (mixed_modes order)

    def ph(freq):  # [C,1] Hz -> [C,n] fp32 phase, reduced in fp64 first
        return torch.remainder(freq.double() * tw.view(1, -1), 2 * torch.pi).float()

    out = torch.empty(n_channels, n, device=device, dtype=torch.int16)
    step = max(1, (1 << 26) // max(n, 1))  # bound the fp64 temporaries
    for r0 in range(0, n_channels, step):
        r1 = min(n_channels, r0 + step)
        s = slice(r0, r1)
        am = A[s] * (1.0 + 0.5 * torch.sin(ph(fm[s]))) * torch.cos(ph(f0 + delta[s]))
        sgn = torch.where(mode[s] == 2, -1.0, 1.0)
        ssb = sum(8000.0 * torch.cos(ph(f0 + sgn * tones[s, k:k + 1])) for k in range(3))
        key = (torch.floor(t * 10.0) % 2 == 0).float().view(1, -1)
        cw = A[s] * key * torch.cos(ph(torch.full_like(delta[s], f0 + 800.0)))
        x = torch.where(mode[s] == 0, am, torch.where(mode[s] == 3, cw, ssb))
        x = x + torch.empty_like(x).uniform_(-32.0, 32.0, generator=g)
        out[s] = torch.clamp(torch.round(x), -32768, 32767).to(torch.int16)
    return out
