#!/bin/bash
mkdir -p gpurun_out/r2
V=minimal-sdr_b200/csrc/variants
for lib in "" $V/libmsdr_ns4.so $V/libmsdr_ns16.so; do
  for ch in 131072; do
    echo "== lib [$lib] c5 $ch"; MSDR_LIBMSDR=$lib timeout 600 python bench.py --config c5 --channels $ch --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | cut -c1-170
    echo "== lib [$lib] ablate 16"; MSDR_LIBMSDR=$lib timeout 600 python bench.py --config c5 --channels $ch --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant 16 2>/dev/null | cut -c1-170
  done
done
