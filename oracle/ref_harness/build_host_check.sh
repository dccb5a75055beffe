#!/bin/bash
# Build oracle/_ref/host_policy_check: the C++ host binding driven by the reference's own AgentModel
# (needs the reference tree, libtorch from the Python environment and libstrikeforce_b200.so).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${SF_REFERENCE:-/root/reference}/StrikeForce-client"
OUT="$ROOT/oracle/_ref"
[ -d "$REF" ] || { echo "reference tree not present; keeping any prebuilt $OUT/host_policy_check" >&2; exit 0; }
mkdir -p "$OUT/build/bots/bot-0.5"
ln -sfn "$REF/bots/bot-0.5/Modules.hpp" "$OUT/build/bots/bot-0.5/Modules.hpp"
TORCH="$(python -c 'import torch, os; print(os.path.dirname(torch.__file__))')"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
g++ -std=c++17 -O1 -w -D_GLIBCXX_USE_CXX11_ABI=1 \
    -I"$OUT/build" -I"$HERE/stubs" -I"$ROOT/include" -I"$ROOT/strikeforce_b200/host" \
    -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include" -I"$CUDA/include" \
    "$HERE/host_policy_check.cpp" -o "$OUT/host_policy_check" \
    -L"$ROOT/strikeforce_b200" -lstrikeforce_b200 -Wl,-rpath,'$ORIGIN/../../strikeforce_b200' \
    -L"$TORCH/lib" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -Wl,-rpath,"$TORCH/lib" -lpthread
echo "built $OUT/host_policy_check"
