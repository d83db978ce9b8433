#!/bin/bash
# Build a copy of the library with extra defines into profiles/ablate/lib_<name>.so (run HERE, no GPU needed), e.g.
#   profiles/build_variant.sh pin0 -DMM_PIN=0        # no constants pinned into vector registers
#   profiles/build_variant.sh pin4 -DMM_PIN=4
#   profiles/build_variant.sh occ2 -DMM_FORCE_BLOCKS=2   # FORCE-only instantiations compiled for two blocks per SM
# and select it on the GPU box with MICMEC_B200_LIB=$PWD/profiles/ablate/lib_<name>.so (r01c_pin_run.sh, r01c_occ2_run.sh).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p profiles/ablate
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --fmad=true \
    --threads 0 "$@" -o profiles/ablate/lib_$name.so micmec_b200/csrc/*.cu
ls -la profiles/ablate/lib_$name.so
