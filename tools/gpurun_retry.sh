#!/bin/bash
# gpurun with retries on "no slot right now" (exit code 3, nothing charged): tools/gpurun_retry.sh [--gpus N] --timeout S -- 'cmd'
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 150 s"
  sleep 150
done
exit 3
