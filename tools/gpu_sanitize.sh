#!/bin/bash
# compute-sanitizer passes over tools/sanitize_smoke.py (memcheck, racecheck, initcheck); summaries under gpurun_out/.
# Usage (repo root, GPU box): bash tools/gpu_sanitize.sh <tag>
TAG=${1:-san}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python tools/sanitize_smoke.py > $OUT/${TAG}_plain.txt 2>&1; echo "plain rc=$?"
for tool in memcheck racecheck initcheck; do
  extra=""
  [ "$tool" = "initcheck" ] && export GMRFB_POOL=0
  SAN_NX=${SAN_NX:-70} timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 \
      python tools/sanitize_smoke.py > $OUT/${TAG}_$tool.txt 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_SMOKE' $OUT/${TAG}_$tool.txt | tr '\n' ' ')"
done
