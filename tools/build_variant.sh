#!/bin/bash
# tools/build_variant.sh NAME FILE.cu "-DX=1 ..."  -> nerfstudio_thermal_b200/lib/variants/NAME.so
# (the default objects of every other translation unit + FILE.cu rebuilt with the extra flags; A/B kernels on the
# GPU box by copying a variant over lib/libtn_b200.so between bench runs)
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
C=$ROOT/nerfstudio_thermal_b200/csrc; L=$ROOT/nerfstudio_thermal_b200/lib
NAME=$1; FILE=$2; EXTRA=${3:-}
mkdir -p $L/variants /tmp/tnv_$NAME
FM=""; case $FILE in tn_encode.cu|tn_geometry.cu|tn_ray.cu|tn_prop.cu|tn_level.cu) FM="-fmad=false";; esac
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I$ROOT/include -Xcompiler -fPIC $FM $EXTRA -c $C/$FILE -o /tmp/tnv_$NAME/${FILE%.cu}.o
OBJS=$(ls $L/obj/*.o | grep -v "/${FILE%.cu}.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $L/variants/$NAME.so $OBJS /tmp/tnv_$NAME/${FILE%.cu}.o -lcudart
echo built $L/variants/$NAME.so
