#!/bin/bash
# round-2 evidence on ONE GPU, final build: bench lines, channel sweep, role counters, launch list, ncu --set full of the chain kernels
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
echo "== bench c5 N=1"; timeout 900 python bench.py --config c5 --steps 5 > $O/c5_n1.json 2> $O/c5_n1.err; cut -c1-160 $O/c5_n1.json; tail -2 $O/c5_n1.err
echo "== bench c3"; timeout 900 python bench.py > $O/c3_n1.json 2> $O/c3_n1.err; cut -c1-160 $O/c3_n1.json; tail -2 $O/c3_n1.err
echo "== bench c4"; timeout 900 python bench.py --config c4 --steps 5 > $O/c4_n1.json 2> $O/c4_n1.err; cut -c1-160 $O/c4_n1.json; tail -2 $O/c4_n1.err
echo "== bench c3 --impl reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/c3_ref.json 2> $O/c3_ref.err; cut -c1-160 $O/c3_ref.json
echo "== channel sweep (32 blocks per update, 128 blocks per step)" | tee $O/channels.txt
for c in 2048 4096 8192 12288 16384 32768 65536 131072 262144 524288 1048576; do for v in 0 8192; do
  echo -n "channels $c variant $v: "
  timeout 300 python bench.py --config c5 --channels $c --steps 3 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s', d['roofline']['kernel'][:14])"
done; done 2>&1 | tee -a $O/channels.txt
echo "== role counters v5 (131072 ch)"; MSDR_PROF=1 timeout 600 python bench.py --config c5 --channels 131072 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 2>&1 >/dev/null | tail -8 | tee $O/role_cycles_v5.txt
echo "== role counters v4 (c3, 138 blocks)"; MSDR_PROF=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --seconds 0.4 --blocks-per-update 138 2>&1 >/dev/null | tail -9 | tee $O/role_cycles_v4.txt
echo "== ablation v5"; for v in 16 32 48 1 2 4 8; do echo -n "variant $v: "; timeout 600 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s')"; done 2>&1 | tee $O/ablation_v5.txt
echo "== launch list (default bench)"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --e2e-steps 0 > $O/launches.log 2>&1
tail -4 $O/launches_bench.csv
echo "== ncu full: c3 (v4), c5 131072 (v5), c4"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 -f -o $O/prof_c3 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --seconds 3 > $O/prof_c3.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 -f -o $O/prof_c5 python bench.py --config c5 --channels 131072 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 > $O/prof_c5.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 -f -o $O/prof_c4 python bench.py --config c4 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 > $O/prof_c4.log 2>&1
ls -la $O/*.ncu-rep
