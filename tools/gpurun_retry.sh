#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>'  -- retries while the pod answers busy (exit 3)
T=$1; G=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i; sleeping 120 s"; sleep 120
done
exit 3
