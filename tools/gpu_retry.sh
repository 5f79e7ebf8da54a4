#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <ngpus> '<command>'   -- retries while the pool answers busy (exit 3)
T=$1; N=$2; shift 2
for i in $(seq 1 40); do
  if [ "$N" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@"; else /usr/local/graft/bin/gpurun --gpus $N --timeout $T -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
