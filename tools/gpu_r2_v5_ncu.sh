#!/bin/bash
# ncu --set full of one launch of the row-block kernel at 131072 channels x 32 blocks (the per-GPU shard of C5 at N=8)
mkdir -p gpurun_out/r2
V=${1:-0}
TAG=${2:-v5}
B="python bench.py --config c5 --channels 131072 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --variant $V"
echo "== pinned chunk test on the row-block kernel"; MSDR_VARIANT=4096 timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "host_update_pipeline" 2>&1 | grep -E "Error|mismatch|passed|failed" | head -6
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 -f -o gpurun_out/r2/prof_$TAG $B > gpurun_out/r2/prof_$TAG.log 2>&1
ls -la gpurun_out/r2/*.ncu-rep; tail -3 gpurun_out/r2/prof_$TAG.log
