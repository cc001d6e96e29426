#!/bin/bash
# memcheck and racecheck over the parity tests that run the time-folded kernel (msdr_chain_v6.cu) and the re-worked row-block epilogue
mkdir -p gpurun_out
run() { tool=$1; shift; echo "== $tool $*"; timeout 900 compute-sanitizer --tool $tool --print-limit 10 python -m pytest "$@" -x -q > gpurun_out/san_tmp.txt 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|out of bounds|hazard" gpurun_out/san_tmp.txt | sort | uniq -c | head -8; cat gpurun_out/san_tmp.txt >> gpurun_out/sanitize_v6.txt; }
rm -f gpurun_out/sanitize_v6.txt
run memcheck tests/test_gpu_chain.py -k "every_kernel_shape and (default or 4096-) or multi_stage or ragged or single_channel or migrat"
run memcheck tests/test_gpu_large.py -k "two_waves"
run racecheck tests/test_gpu_chain.py -k "every_kernel_shape and default"
