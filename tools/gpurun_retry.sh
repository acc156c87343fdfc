#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <gpus> <command...>  — retries while the pod answers "no slot right now" (rc 3)
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [[ "$G" == "1" ]]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"; fi
  rc=$?
  if [[ $rc -ne 3 ]]; then exit $rc; fi
  echo "[retry $i] no slot, sleeping 90 s"; sleep 90
done
exit 3
