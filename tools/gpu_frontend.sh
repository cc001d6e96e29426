#!/bin/bash
# K4 front-end conditioning: parity tests, stand-alone device-resident timing, and the chain with the front end in front of it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend.py -q 2>&1 | tail -2
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/frontend.txt
import sys, torch
sys.path.insert(0, '.')
import minimal_sdr_b200 as m
dev = torch.device('cuda', 0)
st = torch.cuda.Stream()
for C, nb, reps in [(4096, 1024, 3), (65536, 128, 3), (262144, 64, 3)]:
    L = nb * 128
    adc = torch.randint(1800, 2300, (C, L), dtype=torch.int16, device=dev)
    out = torch.empty_like(adc)
    fe = m.Frontend(C)
    fe.set_stream(st.cuda_stream)
    for _ in range(2):
        fe.update_device(adc.data_ptr(), out.data_ptr(), nb, L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fe.update_device(adc.data_ptr(), out.data_ptr(), nb, L)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n = C * L
    print(f"front end  channels={C:7d} blocks={nb:5d}: {ms:8.3f} ms/launch  {n / ms / 1e3:9.0f} Msamples/s  {4 * n / ms / 1e6:7.1f} GB/s algorithmic "
          f"({n / ms / 1e3 * 4 / 6549.4 / 10:.1f} % of the measured HBM roofline)")
    fe.close(); del adc, out
PY
for f in "" "--with-frontend"; do
  echo -n "bench $f: "
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 $f 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), 'ms/step', round(d['value']), 'Msamples/s', d['gpu_launches'], 'launches')"
done 2>&1 | tee -a gpurun_out/frontend.txt
