#!/bin/bash
# compute-sanitizer on the warp-specialised kernels (SURVEY.md section 5 stance): memcheck, racecheck, synccheck
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|hazard" gpurun_out/r02_sanitizer_$tool.log | head -8
done
