#!/bin/bash
# compute-sanitizer over the kernels of the hot path: memcheck, racecheck, synccheck and initcheck on
# tools/sanitize_target.py.  Run on the GPU box:  gpurun --timeout 1800 -- 'bash tools/sanitize.sh'
# The log summaries land in gpurun_out/r2_sanitize_<tool>.txt; copy them to profiles/ when they are clean.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  log=gpurun_out/r2_sanitize_$tool.txt
  timeout ${SANITIZE_TIMEOUT:-1500} compute-sanitizer --tool $tool $extra --print-limit 20 \
      python tools/sanitize_target.py ${SANITIZE_WHAT:-smoke cohort flags stress ingest} > $log.full 2>&1
  echo "exit code $?" >> $log.full
  # keep the summary: every sanitizer line, the target's own last line and the exit code
  grep -E "^=========|sanitize_target|exit code|Error|Traceback" $log.full | head -200 > $log
  tail -4 $log
done
