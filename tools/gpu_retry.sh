#!/bin/bash
# retry a gpurun call while the pod answers busy (exit 3 / transient); usage: gpu_retry.sh <timeout> <script> [gpus]
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  rc=$?
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 120; continue; fi
  exit $rc
done
