#!/bin/bash
echo "== shape test default (v6)"; timeout 300 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "every_kernel_shape and (default or 16384)" 2>&1 | tail -15
echo "== chain file (v6 default)"; timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "not whole_file" 2>&1 | tail -15
