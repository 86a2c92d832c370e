#!/bin/bash
# compute-sanitizer over every kernel family (run under gpurun, one GPU); logs land in gpurun_out/.
# racecheck: shared-memory hazards between the warps of a block (the half-line kernels exchange tiles under
# __syncthreads / named barriers / mbarriers); synccheck: divergent or mismatched barriers; memcheck: bounds.
set -x
python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 python tools/sanitize_cases.py \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|worst|all ok" gpurun_out/sanitize_$tool.log | tail -20
done
