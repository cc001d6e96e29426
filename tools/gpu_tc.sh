#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -15
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/tc_study.txt
import ctypes as C, sys
sys.path.insert(0, '.')
import minimal_sdr_b200 as m
L = m.capi.lib()
for T, rows, nb, kind in [(86, 4096, 64, 1), (102, 4096, 64, 2), (86, 16384, 64, 1), (256, 16384, 64, 1), (256, 16384, 64, 2)]:
    ms = C.c_float(0)
    st = L.msdr_study_fir_demod_tc_time(0, T, rows, nb * 128, kind, 10, C.byref(ms))
    n = rows * nb * 128
    print(f"tensor-core FIR+demod  T={T:3d} rows={rows:5d} blocks={nb} kind={kind}: status {st}  {ms.value:8.3f} ms/launch  {n / ms.value / 1e3:9.0f} Msamples/s")
PY
