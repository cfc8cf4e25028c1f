#!/bin/bash
# gpurun with retries while the pod answers "transient" (busy): [GPUS=N] tools/gpurun_retry.sh <timeout> <command...>
T=$1; shift
G=${GPUS:-1}
for i in $(seq 1 12); do
  if [ "$G" = "1" ]; then
    out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  else
    out=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" 2>&1)
  fi
  echo "$out" | tail -70
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  echo "[retry $i] pod busy, sleeping 150 s"
  sleep 150
done
exit 3
