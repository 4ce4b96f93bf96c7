#!/bin/bash
# A/B builds of single kernels: scripts/build_variant.sh <name> <source.cu> ["-DFOO=1 ..."] [git-rev]
#   -> onssen_b200/libonssen_b200_<name>.so = the current objects with <source> recompiled with the extra defines
#      (or taken from <git-rev>).  Select at run time with ONSSEN_LIB=onssen_b200/libonssen_b200_<name>.so.
set -e
NAME=$1; SRC=$2; DEFS=$3; REV=$4
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
C=$ROOT/onssen_b200/csrc
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I$C -I$ROOT/include"
mkdir -p $C/build/var_$NAME
base=$(basename $SRC .cu)
if [ -n "$REV" ]; then
  git -C $ROOT show $REV:onssen_b200/csrc/$SRC > $C/build/var_$NAME/$SRC
  IN=$C/build/var_$NAME/$SRC
else
  IN=$C/$SRC
fi
$NVCC $FLAGS $DEFS -c $IN -o $C/build/var_$NAME/$base.o
objs=""
for f in capi gemm_tc05 lstm_rec lstm_bwd lstm_bwd_tc pack loss stft extras backward optim kmeans wav; do
  if [ $f == $base ]; then objs="$objs $C/build/var_$NAME/$base.o"; else objs="$objs $C/build/$f.o"; fi
done
$NVCC -shared -o $ROOT/onssen_b200/libonssen_b200_$NAME.so $objs -lcudart
echo "built onssen_b200/libonssen_b200_$NAME.so"
