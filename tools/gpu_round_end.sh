#!/bin/bash
# what the driver runs at round end, in one visit: smoke, GPU tests, reference arm, bench
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference 2>&1 | tail -1 | cut -c1-700
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
