#!/bin/bash
# Builds libquartz_gpu.so in-tree for sm_100a.  -fmad=false: the reference (rustc) never contracts a*b+c.
set -e
cd "$(dirname "$0")/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-Wall -Xptxas -v"
OUT=../libquartz_gpu.so
SRCS="graph.cpp parse.cpp lower.cpp spec.cpp capi.cu interp.cu fused.cu fused_poly.cu spectral.cu"
newest=$(ls -t $SRCS *.h *.cuh ../build.sh | head -1)
if [ -f "$OUT" ] && [ "$OUT" -nt "$newest" ]; then exit 0; fi
mkdir -p ../_obj
pids=()
for s in $SRCS; do
  o=../_obj/${s%.*}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find . -name '*.h' -newer "$o" -o -name '*.cuh' -newer "$o")" ]; then
    ( $NVCC $FLAGS -x cu -c "$s" -o "$o" > ../_obj/${s%.*}.log 2>&1 || { cat ../_obj/${s%.*}.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT ../_obj/*.o -ldl
echo "built $OUT"
