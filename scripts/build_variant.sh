#!/bin/bash
# usage: scripts/build_variant.sh NAME "-DVG_S_DEPTH=8 ..."  -> scripts/variants/NAME.so (tuning experiments only)
set -e
cd "$(dirname "$0")/.."
make -C kangaroo_b200/csrc >/dev/null
O=kangaroo_b200/lib/obj
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2 $2 -c kangaroo_b200/csrc/sgm_fused.cu -o scripts/variants/$1.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scripts/variants/$1.so $O/census.o $O/sgm.o scripts/variants/$1.o $O/wta.o $O/frontback.o $O/median.o $O/engine.o -lcudart
rm scripts/variants/$1.o
