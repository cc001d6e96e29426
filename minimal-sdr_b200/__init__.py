"""minimal-sdr_b200 — B200-native receive DSP chain of FrankBoesing/Minimal-SDR.

Python is test/benchmark orchestration only: `capi` binds the C ABI of csrc/libmsdr.so (include/msdr.h)
with ctypes, `chain` mirrors the reference's object API on top of it, `design` holds the host-side
coefficient designers, `synth` the synthetic IF generators.  All computation happens in hand-written CUDA
kernels (csrc/*.cu, sm_100a); importing `capi` fails loudly if the library has not been built and every compute
call fails if no GPU is present — there is no CPU fallback.

The directory name contains a hyphen, so import it through the `minimal_sdr_b200` shim at the repo root.
"""
from . import anr, capi, chain, design, frontend, shard, syncam, synth, workloads  # noqa: F401
from .capi import MsdrError, lib_path  # noqa: F401
from .chain import ReceiveChain, load_ref_constants  # noqa: F401
from .anr import Anr  # noqa: F401
from .frontend import Frontend  # noqa: F401
from .syncam import SyncAm  # noqa: F401

__all__ = ["anr", "Anr", "capi", "chain", "design", "frontend", "shard", "syncam", "synth", "workloads", "ReceiveChain", "Frontend", "SyncAm", "MsdrError", "lib_path", "load_ref_constants"]
