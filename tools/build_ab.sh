#!/bin/bash
# build_ab.sh NAME [GIT_REV]: build the library from the working tree (or from GIT_REV) into cosim_b200/csrc/_build_ab/lib_NAME.so
# for same-box A/B runs (COSIM_LIB_PATH=... tools/quick_rate.py)
set -e
NAME=$1; REV=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT
if [ -n "$REV" ]; then SRC=/tmp/ab_$NAME; rm -rf $SRC; mkdir -p $SRC; git -C $ROOT archive $REV cosim_b200/csrc include | tar -x -C $SRC; fi
OUT=$ROOT/cosim_b200/csrc/_build_ab; mkdir -p $OUT /tmp/abobj_$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $EXTRA"
rm -f /tmp/abobj_$NAME/*.o
for s in engine engine_gen engine_w24 engine_w12 policy; do
  nvcc $FLAGS -c -o /tmp/abobj_$NAME/$s.o $SRC/cosim_b200/csrc/$s.cu &
done
wait
for s in engine engine_gen engine_w24 engine_w12 policy; do test -f /tmp/abobj_$NAME/$s.o || { echo "build_ab: $s.cu failed"; exit 1; }; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/lib_$NAME.so /tmp/abobj_$NAME/*.o
ls -la $OUT/lib_$NAME.so
