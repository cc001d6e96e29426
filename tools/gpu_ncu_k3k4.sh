#!/bin/bash
# ncu --set full of one launch each of the stand-alone tensor-core FIR + demod kernel (K3) and the front-end kernel (K4)
mkdir -p gpurun_out
cat > /tmp/k3.py <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
import minimal_sdr_b200 as m
L = m.capi.lib()
ms = C.c_float(0)
print(L.msdr_study_fir_demod_tc_time(0, 86, 16384, 64 * 128, 1, 2, C.byref(ms)), ms.value)
PY
cat > /tmp/k4.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import minimal_sdr_b200 as m
dev = torch.device('cuda', 0)
C, nb = 262144, 64
adc = torch.randint(1800, 2300, (C, nb * 128), dtype=torch.int16, device=dev)
out = torch.empty_like(adc)
fe = m.Frontend(C)
for _ in range(3):
    fe.update_device(adc.data_ptr(), out.data_ptr(), nb, nb * 128)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_demod_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_k3 python /tmp/k3.py 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frontend_kernel -s 2 -c 1 -f -o gpurun_out/prof_k4 python /tmp/k4.py 2>&1 | tail -3
for k in k3 k4; do
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$k.ncu-rep --page source --csv > gpurun_out/prof_${k}_src.csv 2>/dev/null
  python tools/ncu_src_summary.py gpurun_out/prof_${k}_src.csv 15 > gpurun_out/prof_${k}_src_summary.txt 2>&1
done
python tools/ncu_metrics_json.py gpurun_out/prof_k3_raw.csv gpurun_out/k3_ncu_metrics.json "ncu --set full --clock-control none: one launch of fir_demod_tc_kernel, 86 taps, 16384 rows x 64 blocks, USB (tools/gpu_ncu_k3k4.sh)"
python tools/ncu_metrics_json.py gpurun_out/prof_k4_raw.csv gpurun_out/k4_ncu_metrics.json "ncu --set full --clock-control none: one launch of frontend_kernel, 262144 channels x 64 blocks (tools/gpu_ncu_k3k4.sh)"
rm -f gpurun_out/prof_k3_src.csv gpurun_out/prof_k4_src.csv
ls -la gpurun_out | tail -12
