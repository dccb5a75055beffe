#!/bin/bash
# Build and run the golden-vector generator of the gradient hub (needs the reference tree and
# libtorch from the Python environment; CPU only).  Writes tests/golden/hub_golden.bin.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${SF_REFERENCE:-/root/reference}/StrikeForce-client"
OUT="$ROOT/oracle/_ref"
[ -d "$REF" ] || { echo "reference tree not present" >&2; exit 0; }
bash "$HERE/build_ref.sh" > /dev/null
TORCH="$(python -c 'import torch, os; print(os.path.dirname(torch.__file__))')"
ln -sfn "$REF/bots/bot-0.5/Modules.hpp" "$OUT/build/bots/bot-0.5/Modules.hpp"
g++ -std=c++17 -O1 -w -D_GLIBCXX_USE_CXX11_ABI=1 \
    -I"$OUT/build" -I"$HERE/stubs" -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include" \
    "$HERE/hub_oracle.cpp" -o "$OUT/hub_oracle" \
    -L"$TORCH/lib" -ltorch -ltorch_cpu -lc10 -Wl,-rpath,"$TORCH/lib" -lpthread
"$OUT/hub_oracle" "$ROOT/tests/golden/hub_golden.bin" 8 3 2
