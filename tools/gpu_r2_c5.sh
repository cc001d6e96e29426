#!/bin/bash
# C5 (2^20 channels, strong scaling) on N GPUs of one box + the default C3 line at the same N (end-to-end attribution per rank)
N=${1:-2}
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1; nproc >> $O/topo_n$N.txt; free -g >> $O/topo_n$N.txt
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"; fi
echo "== c5 N=$N"; timeout 1200 $L bench.py --gpus $N --config c5 --steps 5 --no-cpu > $O/c5_n$N.json 2> $O/c5_n$N.err; cut -c1-250 $O/c5_n$N.json; tail -3 $O/c5_n$N.err
echo "== c3 N=$N"; timeout 1200 $L bench.py --gpus $N --steps 5 --no-cpu > $O/c3_n$N.json 2> $O/c3_n$N.err; cut -c1-250 $O/c3_n$N.json; tail -3 $O/c3_n$N.err
