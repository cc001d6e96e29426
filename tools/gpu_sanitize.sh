#!/bin/bash
# compute-sanitizer over the smoke case (40 channels x 12 blocks, three updates) and a 256-tap / SYNCAM / ANR mix: memcheck, then racecheck
mkdir -p gpurun_out
for tool in ${TOOLS:-memcheck racecheck}; do
  echo "== $tool"
  timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.txt 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/sanitize_$tool.txt | sort | uniq -c | sort -rn | head -12
done
