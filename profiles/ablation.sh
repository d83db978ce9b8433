#!/bin/bash
# Build ablated copies of the library (MM_ABLATE bits, see mm_march.cuh) into profiles/ablate/ - run HERE (no GPU needed).
# usage: profiles/ablation.sh "0 1 2 4 6 ..."
set -e
cd "$(dirname "$0")/.."
for a in ${1:-"0 1 2 4 8 16 32"}; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --fmad=true \
      -DMM_ABLATE=$a -o profiles/ablate/lib_$a.so micmec_b200/csrc/*.cu &
done
wait
ls -la profiles/ablate/
