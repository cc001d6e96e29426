#!/bin/bash
mkdir -p gpurun_out/r2p
echo "== role counters v4 at the bench shape (4096 ch x 1024 blocks per update)"
MSDR_PROF=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 2>&1 >/dev/null | tail -40 | tee gpurun_out/r2p/role_cycles_v4_1024.txt
