#!/bin/bash
# compute-sanitizer passes over the hot path: memcheck, racecheck, synccheck, initcheck on a tiny workload
# (tools/sanitize_target.py); logs + one-line summaries into gpurun_out/ (copied to profiles/ per round).
# Usage (GPU box): bash tools/run_sanitizers.sh [tools...]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=("$@"); [ ${#TOOLS[@]} -eq 0 ] && TOOLS=(memcheck racecheck synccheck)
SAN=/usr/local/cuda/bin/compute-sanitizer
for t in "${TOOLS[@]}"; do
  log=gpurun_out/sanitizer_${t}.log
  echo "== compute-sanitizer --tool $t" | tee "$log"
  start=$(date +%s)
  timeout "${SAN_TIMEOUT:-900}" "$SAN" --tool "$t" --print-limit 20 --error-exitcode 66 \
      python tools/sanitize_target.py >> "$log" 2>&1
  rc=$?
  end=$(date +%s)
  summary=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
  echo "tool=$t exit=$rc seconds=$((end-start)) ${summary}" | tee -a gpurun_out/sanitizer_summary.txt
  grep -E "\[sanitize\]" "$log" | tail -5
done
