timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_large.py tests/test_gpu_stages.py -x -q 2>&1 | tail -2
for a in "--variant 0" "--variant 128" "--variant 144" "--variant 160" "--variant 176"; do echo -n "$a: "; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 $a 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s')"; done
VARIANTS="128" bash tools/gpu_v4_prof.sh
