#!/bin/bash
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
echo "== bench c5 N=1"; timeout 900 python bench.py --config c5 --steps 5 > $O/c5_n1.json 2> $O/c5_n1.err; cut -c1-160 $O/c5_n1.json; tail -1 $O/c5_n1.err
echo "== channel sweep (32 blocks per update, 128 blocks per step)" | tee $O/channels.txt
for c in 12288 16384 32768 65536 131072 262144 524288 1048576; do for v in 0 8192; do
  echo -n "channels $c variant $v: "
  timeout 300 python bench.py --config c5 --channels $c --steps 3 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s', d['roofline']['kernel'][:14])"
done; done 2>&1 | tee -a $O/channels.txt
echo "== role counters v5 (131072 ch)"; MSDR_PROF=1 timeout 600 python bench.py --config c5 --channels 131072 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 2>&1 >/dev/null | tail -8 | tee $O/role_cycles_v5.txt
echo "== ablation v5"; for v in 16 32 48 1 2 4 8 12; do echo -n "variant $v: "; timeout 600 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s')"; done 2>&1 | tee $O/ablation_v5.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 -f -o $O/prof_c5 python bench.py --config c5 --channels 131072 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 > $O/prof_c5.log 2>&1
ls -la $O/*.ncu-rep
