#!/bin/bash
# compute-sanitizer passes over one small head step (the mbarrier / TMEM kernels need it)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_step.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|loss|Error|error" gpurun_out/r02_sanitizer_$tool.log | head -8
done
