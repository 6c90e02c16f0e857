#!/bin/bash
# compute-sanitizer over tools/sanitize_smoke.py, one tool at a time; summary lines into gpurun_out/<tag>_sanitizer.txt
tag=${1:-r2b}
out=gpurun_out/${tag}_sanitizer.txt
: > $out
for tool in memcheck synccheck initcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/${tag}_san_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_san_$tool.log | tail -1)" >> $out
  grep -c "finite True" gpurun_out/${tag}_san_$tool.log >> $out
done
cat $out
