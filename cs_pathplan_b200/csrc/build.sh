#!/bin/sh
# Build cs_pathplan_b200/libmsnap_b200.so (sm_100a only, in-tree so it travels to the GPU box).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libmsnap_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false ${MSNAP_EXTRA} \
    -Xcompiler -fPIC,-O2,-Wall -shared -cudart static \
    ${MSNAP_PTXAS_V:+-Xptxas -v} \
    -o "$OUT" "$HERE/msnap_capi.cu"
echo "built $OUT"
