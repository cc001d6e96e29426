#!/bin/bash
# K5 LMS notch / noise reduction: parity tests and stand-alone device-resident timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_anr.py -q 2>&1 | tail -3
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/anr.txt
import sys, torch
sys.path.insert(0, '.')
import minimal_sdr_b200 as m
dev = torch.device('cuda', 0)
st = torch.cuda.Stream()
for C, nb, reps in [(4096, 64, 2), (65536, 16, 2)]:
    L = nb * 128
    a = torch.randint(-3000, 3000, (C, L), dtype=torch.int16, device=dev)
    anr = m.Anr(C)
    anr.set_stream(st.cuda_stream)
    anr.update_device(1, a.data_ptr(), nb, L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        anr.update_device(1, a.data_ptr(), nb, L)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n = C * L
    print(f"ANR  channels={C:6d} blocks={nb:4d}: {ms:9.3f} ms/launch  {n / ms / 1e3:8.0f} Msamples/s  ({ms * 1e-3 * 1.965e9 / L:6.0f} cycles per sample-step per warp)")
    anr.close()
PY
