#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe). Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0"
echo "== launch list"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1
tail -3 gpurun_out/launches.csv
echo "== full set, chain kernel"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 -f -o gpurun_out/prof_chain $B --seconds 3 > gpurun_out/prof.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/prof.log
