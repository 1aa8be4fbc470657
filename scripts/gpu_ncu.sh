#!/bin/bash
# usage (on the GPU box): scripts/gpu_ncu.sh NAME SKIP COUNT KREGEX <bench args...>
# ncu --set full of COUNT launches matching KREGEX after SKIP matching launches; the report stays on the box (too big for
# gpurun_out), only its raw-page CSV and (gzipped) source-page CSV come back.
NAME=$1; SKIP=$2; COUNT=$3; KRE=$4; shift 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $COUNT -f -o /tmp/$NAME python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${NAME}_ncu.log 2>&1
ncu -i /tmp/$NAME.ncu-rep --page raw --csv > gpurun_out/${NAME}_raw.csv 2>/dev/null
ncu -i /tmp/$NAME.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${NAME}_source.csv.gz
ls -la /tmp/$NAME.ncu-rep gpurun_out/${NAME}_*
