#!/bin/bash
# usage: scripts/gpurun_retry.sh OUTFILE TIMEOUT cmd...   (retries while the pod answers busy)
OUT=$1; TO=$2; shift 2
for i in $(seq 1 30); do
  gpurun --timeout $TO -- "$@" > $OUT 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
