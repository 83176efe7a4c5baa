#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3 / "transient"): tools/gpu_retry.sh <timeout> '<command>'
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit $rc
done
echo "gpu_retry: no slot after 40 attempts"; exit 3
