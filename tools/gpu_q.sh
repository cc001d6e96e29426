timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash tools/gpu_tc.sh | tail -5
for a in "--variant 0" "--variant 32" "--channels 65536 --seconds 0.3715 --blocks-per-update 128 --variant 0" "--channels 65536 --seconds 0.3715 --blocks-per-update 128 --variant 32"; do echo -n "$a: "; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 $a 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s')"; done
