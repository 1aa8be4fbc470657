#!/bin/bash
# usage: scripts/build_variant.sh NAME "-DVG_S_DEPTH=8 ..." [source.cu ...]  -> scripts/variants/NAME.so
# Tuning experiments only: recompiles the listed sources (default sgm_fused.cu) with the extra flags and links them
# with the objects of the regular build.
set -e
cd "$(dirname "$0")/.."
make -C kangaroo_b200/csrc -j8 >/dev/null
O=kangaroo_b200/lib/obj
NAME=$1; FLAGS=$2; shift 2 || true
SRCS=${@:-sgm_fused.cu}
mkdir -p scripts/variants/obj_$NAME
OBJS=""
for s in census sgm sgm_hsweep_dispatch sgm_fused wta frontback median volfilter engine split_engine; do
  if [[ " $SRCS " == *" $s.cu "* ]]; then
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2 $FLAGS -c kangaroo_b200/csrc/$s.cu -o scripts/variants/obj_$NAME/$s.o
    OBJS="$OBJS scripts/variants/obj_$NAME/$s.o"
  else
    OBJS="$OBJS $O/$s.o"
  fi
done
for p in 1 2 4 8 16; do
  if [[ " $SRCS " == *" sgm_hsweep.cu "* ]]; then
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2 $FLAGS -DHS_PART=$p -c kangaroo_b200/csrc/sgm_hsweep.cu -o scripts/variants/obj_$NAME/sgm_hsweep_dpl$p.o
    OBJS="$OBJS scripts/variants/obj_$NAME/sgm_hsweep_dpl$p.o"
  else
    OBJS="$OBJS $O/sgm_hsweep_dpl$p.o"
  fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scripts/variants/$NAME.so $OBJS -lcudart
rm -rf scripts/variants/obj_$NAME
