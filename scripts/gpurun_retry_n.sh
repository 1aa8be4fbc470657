#!/bin/bash
# usage: scripts/gpurun_retry_n.sh NGPUS OUTFILE TIMEOUT cmd...   (retries while the pod answers busy)
N=$1; OUT=$2; TO=$3; shift 3
for i in $(seq 1 40); do
  gpurun --gpus $N --timeout $TO -- "$@" > $OUT 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
