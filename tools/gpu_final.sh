#!/bin/bash
# Round-end evidence: bench line, ncu launch list + full capture of the chain kernel, text summaries for profiles/.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for v in 16 32 48 2048 256 64; do timeout 300 python bench.py --steps 3 --warmup 3 --variant $v --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ablation variant $v', round(d['value']), 'Msamples/s', round(d['ms_per_step'],2), 'ms/step')" | tee -a gpurun_out/ablation.txt; done
bash tools/gpu_profile.sh
ncu -i gpurun_out/prof_chain.ncu-rep --page raw --csv > gpurun_out/prof_chain_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_chain.ncu-rep --page source --csv > gpurun_out/prof_chain_src.csv 2>/dev/null
python tools/ncu_src_summary.py gpurun_out/prof_chain_src.csv 25 > gpurun_out/prof_chain_src_summary.txt 2>&1
python tools/ncu_metrics_json.py gpurun_out/prof_chain_raw.csv gpurun_out/chain_kernel_ncu_metrics.json "ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --seconds 3 (tools/gpu_profile.sh): one launch of 4096 channels x 1024 blocks"
