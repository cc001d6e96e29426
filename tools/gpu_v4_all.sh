#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
MSDR_VARIANT=256 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
MSDR_VARIANT=128 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
MSDR_VARIANT=64 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_v4_ablate.sh
VARIANTS="0" bash tools/gpu_v4_prof.sh | head -12
