#!/bin/bash
# Build the C++ host of the batched tick (b200_play) against libtorch from the Python environment and
# libstrikeforce_b200.so.  In-tree: the binary travels with the repository snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
TORCH="$(python -c 'import torch, os; print(os.path.dirname(torch.__file__))')"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
g++ -std=c++17 -O2 -w -D_GLIBCXX_USE_CXX11_ABI=1 \
    -I"$ROOT/include" -I"$HERE" -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include" -I"$CUDA/include" \
    "$HERE/b200_play.cpp" -o "$HERE/b200_play" \
    -L"$ROOT/strikeforce_b200" -lstrikeforce_b200 -Wl,-rpath,'$ORIGIN/..' \
    -L"$TORCH/lib" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -Wl,-rpath,"$TORCH/lib" -lpthread
echo "built $HERE/b200_play"
