#!/bin/bash
# TEST INFRASTRUCTURE — compile the real kernel sources against the CPU emulator (no GPU needed).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
SRC="$ROOT/multiview_motion_capture_b200/csrc"
OUT="$HERE/libmvmc_emu.so"
FLAGS="-O1 -g -fPIC -std=c++17 -DMVMC_EMU -ffp-contract=off -I$ROOT/include -I$HERE -I$SRC -Wno-unused-variable"
objs=""
pids=""
for f in affinity als assign ik ingest matchers pipeline; do
  rm -f "$HERE/$f.emu.o"
  g++ $FLAGS -x c++ -c "$SRC/$f.cu" -o "$HERE/$f.emu.o" &
  pids="$pids $!"
  objs="$objs $HERE/$f.emu.o"
done
rm -f "$HERE/als_small.emu.o"
g++ $FLAGS -DAL_VARIANT=small -DAL_FM_=3 -DAL_THREADS_=128 -x c++ -c "$SRC/als.cu" -o "$HERE/als_small.emu.o" &
pids="$pids $!"
objs="$objs $HERE/als_small.emu.o"
g++ $FLAGS -c "$HERE/emu_main.cpp" -o "$HERE/emu_main.emu.o" &
pids="$pids $!"
for p in $pids; do wait $p; done   # (a bare `wait` would swallow a failed compile and relink the stale object)
g++ -shared -o "$OUT" $objs "$HERE/emu_main.emu.o" -lpthread
echo "built $OUT"
