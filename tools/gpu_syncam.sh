#!/bin/bash
# K6 synchronous-AM PLL: tolerance tests and stand-alone device-resident timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_syncam.py -q 2>&1 | tail -6
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/syncam.txt
import sys, ctypes as C, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import minimal_sdr_b200 as m
import syncam_lib as sl
I, Q = sl.baseband(64, 128 * 200, seed=1)
yg = m.SyncAm(64).update(I, Q).astype(np.float64); yo = sl.OrcSyncAm(64).run(I, Q).astype(np.float64)
print(f"parity vs oracle: max |diff| {np.abs(yg - yo).max():.0f} LSB, differing samples {100 * (yg != yo).mean():.3f} %, relative RMS {np.sqrt(np.mean((yg - yo) ** 2)) / np.sqrt(np.mean(yo ** 2)):.3g}")
L = m.capi.lib()
dev = torch.device('cuda', 0)
st = torch.cuda.Stream()
for Cn, nb, reps in [(4096, 64, 2), (65536, 16, 2)]:
    n = nb * 128
    a = torch.randint(-3000, 3000, (Cn, n), dtype=torch.int16, device=dev); b = torch.randint(-3000, 3000, (Cn, n), dtype=torch.int16, device=dev); o = torch.empty_like(a)
    sc = m.SyncAm(Cn)
    L.msdr_syncam_set_stream(sc.h, C.c_void_p(st.cuda_stream))
    call = lambda: L.msdr_syncam_update_device(sc.h, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(o.data_ptr()), nb, n)
    call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): call()
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"SYNCAM  channels={Cn:6d} blocks={nb:4d}: {ms:9.3f} ms/launch  {Cn * n / ms / 1e3:8.0f} Msamples/s")
    sc.close()
PY
