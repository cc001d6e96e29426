#!/bin/bash
# whole GPU suite + the three bench configurations on one GPU
mkdir -p gpurun_out/r2
O=gpurun_out/r2
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest2.log
echo "== bench c5 N=1"; timeout 900 python bench.py --config c5 --steps 5 --no-cpu > $O/c5_n1.json 2> $O/c5_n1.err; cut -c1-200 $O/c5_n1.json; tail -3 $O/c5_n1.err
echo "== bench c3"; timeout 900 python bench.py > $O/c3_n1.json 2> $O/c3_n1.err; cut -c1-200 $O/c3_n1.json; tail -3 $O/c3_n1.err
