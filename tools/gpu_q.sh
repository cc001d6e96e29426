mkdir -p gpurun_out; rm -f gpurun_out/ablation.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200
bash tools/gpu_final.sh 2>&1 | tail -25
VARIANTS="0 16 32" bash tools/gpu_v4_prof.sh > /dev/null 2>&1; cat gpurun_out/v4_prof.txt
