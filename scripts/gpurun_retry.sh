#!/bin/bash
# usage: scripts/gpurun_retry.sh [gpurun options] -- '<command>' : retries while the pod answers "busy" (exit 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $i answered busy; sleeping 120 s"; sleep 120
done
exit 3
