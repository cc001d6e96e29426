#!/bin/bash
# round 2, first visit: the whole GPU suite (incl. the new at-size tests), the three bench configurations, host topology
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu | head -40 >> $O/topo.txt; cat /sys/devices/system/node/node*/cpulist >> $O/topo.txt 2>&1; nproc >> $O/topo.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest.log
echo "== bench c3"; timeout 900 python bench.py --steps 5 > $O/bench_c3.json 2> $O/bench_c3.err; cut -c1-300 $O/bench_c3.json; tail -3 $O/bench_c3.err
echo "== bench c4"; timeout 900 python bench.py --config c4 --steps 3 --no-cpu > $O/bench_c4.json 2> $O/bench_c4.err; cut -c1-300 $O/bench_c4.json; tail -3 $O/bench_c4.err
echo "== bench c5"; timeout 900 python bench.py --config c5 --steps 3 --no-cpu > $O/bench_c5.json 2> $O/bench_c5.err; cut -c1-300 $O/bench_c5.json; tail -3 $O/bench_c5.err
echo "== bench c5 65536 ch"; timeout 900 python bench.py --config c5 --channels 65536 --steps 5 --no-cpu --e2e-steps 0 > $O/bench_c5_64k.json 2> $O/bench_c5_64k.err; cut -c1-300 $O/bench_c5_64k.json; tail -3 $O/bench_c5_64k.err
