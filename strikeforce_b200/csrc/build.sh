#!/bin/bash
# Build libstrikeforce_b200.so (CUDA kernels + C ABI) in-tree for sm_100a.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="${SF_OUT:-$ROOT/strikeforce_b200/libstrikeforce_b200.so}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-Wall,-Wno-unused-parameter -shared \
    -I"$ROOT/include" -I"$HERE" ${SF_NVCC_EXTRA:-} \
    "$HERE/sf_lib.cu" -o "$OUT"
echo "built $OUT"
