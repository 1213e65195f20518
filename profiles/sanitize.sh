#!/bin/bash
# compute-sanitizer evidence for the hand-rolled mbarrier / TMEM / TMA pipelines:
#   bash profiles/sanitize.sh [outdir]      -> <outdir>/sanitize_{memcheck,racecheck,synccheck,initcheck}.log
# Each tool runs profiles/sanitize_targets.py (small shapes, results checked against torch) under its own timeout.
out=${1:-gpurun_out}
mkdir -p "$out"
for tool in memcheck synccheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_targets.py > "$out/sanitize_$tool.log" 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$out/sanitize_$tool.log" | tail -1)"
done
