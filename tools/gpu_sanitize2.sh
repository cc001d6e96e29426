#!/bin/bash
# memcheck over the parity tests that cover every kernel shape and every kernel
mkdir -p gpurun_out
run() { echo "== $*"; timeout 700 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest "$@" -x -q > gpurun_out/san_tmp.txt 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/san_tmp.txt | sort | uniq -c | head -8; cat gpurun_out/san_tmp.txt >> gpurun_out/sanitize_memcheck_tests.txt; }
rm -f gpurun_out/sanitize_memcheck_tests.txt
run tests/test_gpu_chain.py -k "every_kernel_shape or long_taps or multi_stage or syncam or anr or adversarial"
run tests/test_gpu_large.py -k "two_waves"
run tests/test_gpu_frontend.py tests/test_gpu_anr.py tests/test_gpu_syncam.py
run tests/test_gpu_tc.py -k "not exhaustive and not sqrt"
