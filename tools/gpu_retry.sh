#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <script> [gpus]  -- retries while the pool answers busy (exit 3)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then gpurun --timeout $T -- "bash $S"; else gpurun --gpus $G --timeout $T -- "bash $S"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
