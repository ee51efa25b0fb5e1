#!/bin/bash
# Stage-geometry sweep of the channels-last TMA backward (run on the GPU box): recompiles whitening_cl_tma.cu with -D overrides.
# usage: tools/cl_sweep.sh "<box> <stages> <relu_box> <relu_stages>" ...
cd "$(dirname "$0")/.."
PKG=wt-pse-code_b200
for cfg in "$@"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DWTPSE_CL_APPLY_BOX=$1 -DWTPSE_CL_APPLY_STAGES=$2 \
       -DWTPSE_CL_APPLY_RELU_BOX=$3 -DWTPSE_CL_APPLY_RELU_STAGES=$4 $EXTRA -c $PKG/csrc/whitening_cl_tma.cu -o $PKG/build/whitening_cl_tma.o || exit 1
  nvcc -shared -o $PKG/libwtpse_b200.so $PKG/build/*.o || exit 1
  echo "== box=$1 stages=$2 relu_box=$3 relu_stages=$4 $EXTRA"
  PROBE_TIMES=1 python tools/loss_probe.py 2 cl,cl_relu 2>&1 | grep pair
done
