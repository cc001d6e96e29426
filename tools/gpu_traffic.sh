#!/bin/bash
# quick DRAM-traffic check of one long launch (ncu, a few metrics only) + the bench value
B="python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --seconds 3"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:chain_kernel -c 1 --csv --log-file gpurun_out/traffic.csv $B > /dev/null 2>&1
grep -v "^==" gpurun_out/traffic.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'ms/step', round(d['value']), 'Msamples/s', d['clocks'])"
