#!/bin/bash
# tensor-core chain kernel (variant 64): full GPU parity suite, then the ablation table beside the CUDA-core kernel
mkdir -p gpurun_out
MSDR_VARIANT=64 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
bash tools/gpu_v4_ablate.sh
