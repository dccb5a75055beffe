#!/bin/bash
# Build the reference tick engine, unmodified, into oracle/_ref/libsfref.so.
#
# Nothing is copied from the reference's SOURCES: oracle/_ref/build/ is a farm of
# symlinks to the headers where they lie under $SF_REFERENCE (default /root/reference),
# plus two generated one-line redirect headers (selected_agent.hpp -> the oracle Agent,
# selected_custom.hpp -> the reference's own bots/bot-0.5/Custom.hpp), which is exactly
# the compile-time plugin mechanism the reference documents (README.md:288-300).
# The reference's DATA files (map/, Items/, character/, the test account sheet) are
# copied into oracle/_ref/rundir/ because the engine opens them by relative path at run
# time and /root/reference does not exist on the GPU box.  oracle/_ref/ is git-ignored.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${SF_REFERENCE:-/root/reference}/StrikeForce-client"
OUT="$ROOT/oracle/_ref"
if [ ! -d "$REF" ]; then
    echo "build_ref.sh: $REF not present; keeping any prebuilt $OUT/libsfref.so" >&2
    exit 0
fi
mkdir -p "$OUT/build/bots/bot-0.5" "$OUT/build/bots/bot-oracle" "$OUT/rundir"
for h in gameplay.hpp Character.hpp Item.hpp random.hpp basic.hpp GraphicPrinter.hpp macros.hpp inet_for_windows.hpp; do
    ln -sfn "$REF/$h" "$OUT/build/$h"
done
ln -sfn "$REF/bots/bot-0.5/Custom.hpp" "$OUT/build/bots/bot-0.5/Custom.hpp"
ln -sfn "$HERE/bot-oracle/Agent.hpp" "$OUT/build/bots/bot-oracle/Agent.hpp"
echo '#include "./bots/bot-oracle/Agent.hpp"' > "$OUT/build/selected_agent.hpp"
echo '#include "bots/bot-0.5/Custom.hpp"'     > "$OUT/build/selected_custom.hpp"
# run-time data
rm -rf "$OUT/rundir/map" "$OUT/rundir/Items" "$OUT/rundir/character"
cp -r "$REF/map" "$REF/Items" "$REF/character" "$OUT/rundir/"
cp -f "$REF/accounts/game/1/info, 1.txt" "$OUT/rundir/player_account1.txt"
mkdir -p "$OUT/rundir/accounts/game/1"   # give_info() announces this file in a live online match
cp -f "$REF/accounts/game/1/info, 1.txt" "$OUT/rundir/accounts/game/1/info, 1.txt"
g++ -std=c++17 -O2 -fPIC -shared -w \
    -I"$OUT/build" -I"$HERE/stubs" -I"$ROOT/include" \
    "$HERE/harness.cpp" -o "$OUT/libsfref.so" -lpthread
# the same engine without the harness's capacity / out-of-bounds checks: bench.py times both and reports
# what the checks cost (they are the harness's, not the reference's)
g++ -std=c++17 -O2 -fPIC -shared -w -DSFREF_NO_GUARDS \
    -I"$OUT/build" -I"$HERE/stubs" -I"$ROOT/include" \
    "$HERE/harness.cpp" -o "$OUT/libsfref_noguard.so" -lpthread
echo "built $OUT/libsfref.so"
