#!/bin/bash
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
echo "== bench c4"; timeout 900 python bench.py --config c4 --steps 5 > $O/c4_n1.json 2> $O/c4_n1.err; cut -c1-160 $O/c4_n1.json; tail -2 $O/c4_n1.err
echo "== role counters c4"; MSDR_PROF=1 timeout 600 python bench.py --config c4 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 2>&1 >/dev/null | tail -8 | tee $O/role_cycles_c4.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 -f -o $O/prof_c4 python bench.py --config c4 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 > $O/prof_c4.log 2>&1
ls -la $O/prof_c4.ncu-rep
