#!/bin/bash
# Submit a gpurun call and keep retrying while the pod answers "transient" (no slot free; nothing is charged).
#   tools/gpu/submit.sh <timeout_s> <gpus> '<command>'
t=$1; g=$2; shift 2
for attempt in $(seq 1 40); do
  if [ "$g" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$@" 2>&1); fi
  if echo "$out" | grep -q "status=transient"; then echo "[submit] attempt $attempt: transient, retrying in 90 s"; sleep 90; continue; fi
  echo "$out"
  exit 0
done
echo "[submit] gave up"
exit 3
