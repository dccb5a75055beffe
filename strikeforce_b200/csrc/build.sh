#!/bin/bash
# Build libstrikeforce_b200.so (CUDA kernels + C ABI) in-tree for sm_100a.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
# SF_GEOMETRY=ROWSxCOLS (e.g. 40x128) builds the library for a larger arena next to the default one
GEO=""
TAG=""
if [ -n "${SF_GEOMETRY:-}" ] && [ "${SF_GEOMETRY}" != "30x100" ]; then
    GEO="-DSF_ROWS=${SF_GEOMETRY%x*} -DSF_COLS=${SF_GEOMETRY#*x}"
    TAG="_${SF_GEOMETRY}"
fi
OUT="${SF_OUT:-$ROOT/strikeforce_b200/libstrikeforce_b200${TAG}.so}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-Wall,-Wno-unused-parameter -shared \
    -I"$ROOT/include" -I"$HERE" $GEO ${SF_NVCC_EXTRA:-} \
    "$HERE/sf_lib.cu" -o "$OUT"
echo "built $OUT"
