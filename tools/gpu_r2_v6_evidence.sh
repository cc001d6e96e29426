#!/bin/bash
# round-2 evidence for the time-folded kernel (msdr_chain_v6.cu) on ONE GPU: tests, bench lines, small-channel sweep against the chain
# kernel, role counters, ablation, launch list, ncu --set full
mkdir -p gpurun_out/r2v6
O=gpurun_out/r2v6
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench c3"; timeout 900 python bench.py > $O/c3_n1.json 2> $O/c3_n1.err; cut -c1-200 $O/c3_n1.json; tail -2 $O/c3_n1.err
echo "== bench c3 --with-frontend"; timeout 900 python bench.py --with-frontend --no-cpu --e2e-steps 0 > $O/c3_fe.json 2> $O/c3_fe.err; cut -c1-200 $O/c3_fe.json
echo "== channel sweep, 1024 blocks per update (C3 shape): time-folded kernel (0) against the chain kernel (16384)" | tee $O/channels_small.txt
for c in 1024 2048 3072 4096 4736 6144 8192; do for v in 0 16384; do
  echo -n "channels $c variant $v: "
  timeout 300 python bench.py --channels $c --seconds 2.97 --steps 5 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s', d['roofline']['kernel'][:14])"
done; done 2>&1 | tee -a $O/channels_small.txt
echo "== role counters v6 (c3, 1024 blocks)"; MSDR_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 --seconds 2.97 2>&1 >/dev/null | tail -11 | tee $O/role_cycles_v6.txt
echo "== ablation v6"; for v in 16 32 48 2; do echo -n "variant $v: "; timeout 600 python bench.py --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s')"; done 2>&1 | tee $O/ablation_v6.txt
for a in 4 8 12 16; do echo -n "MSDR_ABLATE $a: "; MSDR_ABLATE=$a timeout 600 python bench.py --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s')"; done 2>&1 | tee -a $O/ablation_v6.txt
echo "== launch list (default bench)"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --e2e-steps 0 > $O/launches.log 2>&1
tail -4 $O/launches_bench.csv
echo "== ncu full: c3 (v6)"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 -f -o $O/prof_c3 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --seconds 3 > $O/prof_c3.log 2>&1
ls -la $O/*.ncu-rep
