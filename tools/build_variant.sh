#!/bin/bash
# tools/build_variant.sh NAME SED_EXPR [FILE]: an A/B build of libgtp_b200.so with one source file edited by sed,
# written to build/ab/libgtp_b200_NAME.so (load it with GTP_B200_LIB=...; bench.py prints the path and sha it timed).
set -e
NAME=$1; EXPR=$2; FILE=${3:-gtp_fast_f32.cu}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/python-geotiepoints_b200/geotiepoints_b200/csrc
OUT=$ROOT/build/ab; mkdir -p $OUT/$NAME
cp $CSRC/*.cu $CSRC/*.cuh $CSRC/*.h $OUT/$NAME/
sed -i -E "$EXPR" $OUT/$NAME/$FILE
cd $OUT/$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$ROOT/include"
sed -i 's#../../../include/gtp_b200.h#gtp_b200.h#' *.cu *.cuh *.h 2>/dev/null || true
# every translation unit of csrc/ (the Makefile's flags: no FMA contraction in the rounding-faithful ones)
for f in gtp_generic gtp_fast_f32 gtp_fast5_f32 gtp_legacy gtp_vii gtp_grid; do nvcc $FLAGS -fmad=false -c $f.cu -o $f.o & done
for f in gtp_fast gtp_fast5 gtp_fast_xyz gtp_capi; do nvcc $FLAGS -c $f.cu -o $f.o & done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libgtp_b200_$NAME.so *.o
echo $OUT/libgtp_b200_$NAME.so
