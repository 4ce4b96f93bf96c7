#!/bin/bash
# Build libonssen_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libonssen_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${ONSSEN_DEFS} -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p build
pids=()
for f in capi gemm_tc05 lstm_rec lstm_bwd lstm_bwd_tc pack loss stft extras backward optim kmeans wav; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ tc05.cuh -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ lstm_bwd.cuh -nt build/$f.o ] || [ ../../include/onssen_b200.h -nt build/$f.o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT build/capi.o build/gemm_tc05.o build/lstm_rec.o build/pack.o build/loss.o build/stft.o build/extras.o build/backward.o build/lstm_bwd.o build/lstm_bwd_tc.o build/optim.o build/kmeans.o build/wav.o -lcudart
echo "built $OUT"
