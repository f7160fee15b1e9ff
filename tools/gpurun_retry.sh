#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout_s> '<command>'  — retries while the pod answers "busy" (exit code 3)
T=$1; shift
G=${GPUS:-1}
for i in $(seq 1 30); do
  if [ "$G" -gt 1 ]; then
    /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"
  else
    /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
