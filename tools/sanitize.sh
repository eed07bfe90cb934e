#!/bin/bash
# compute-sanitizer over tools/sanitize_case.py (one GPU).  usage: tools/sanitize.sh <tag>
T=${1:-x}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  n=5; [ $tool = racecheck ] && n=2
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py $n > gpurun_out/sanitize_${tool}_$T.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok:" gpurun_out/sanitize_${tool}_$T.log | tail -5
done
