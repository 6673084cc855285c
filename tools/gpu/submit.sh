#!/bin/bash
# Submit a gpurun call from a FROZEN copy of the working tree and keep retrying while the pod answers
# "transient" (no slot free; nothing is charged).  gpurun snapshots /root/repo when a box becomes free, which
# can be minutes after the submission; the frozen tarball (sources + the built library, consistent with each
# other) makes the run independent of edits made while the call is queued.
#   tools/gpu/submit.sh <timeout_s> <gpus> '<command run from the root of the frozen copy>'
t=$1; g=$2; shift 2
cd "$(dirname "$0")/../.."
mkdir -p .frozen
rm -f .frozen/*.tar
tag=$(date +%s)
tar --exclude=./.git --exclude=./gpurun_out --exclude=./build --exclude=./.frozen --exclude=__pycache__ --exclude=./.pytest_cache \
    -cf .frozen/$tag.tar .
cmd="rm -rf /tmp/fz && mkdir -p /tmp/fz && tar -xf .frozen/$tag.tar -C /tmp/fz && mkdir -p gpurun_out && ln -sfn \$PWD/gpurun_out /tmp/fz/gpurun_out && cd /tmp/fz && $*"
for attempt in $(seq 1 60); do
  if [ "$g" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$cmd" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$cmd" 2>&1); fi
  if echo "$out" | grep -q "status=transient"; then echo "[submit] attempt $attempt: transient, retrying in 60 s"; sleep 60; continue; fi
  echo "$out"
  rm -f .frozen/$tag.tar
  exit 0
done
echo "[submit] gave up"
rm -f .frozen/$tag.tar
exit 3
