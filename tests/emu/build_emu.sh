#!/bin/sh
# Build the UNIT-TEST double of the device shim (see emu_shim.cpp) together with
# the real host layer into tests/_emu/libfftw3_b200_emu.so.  Test-only artefact.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT=$ROOT/tests/_emu
mkdir -p "$OUT"
H=$ROOT/fftw3_b200/csrc/host
CF="-O2 -fPIC -std=gnu11 -I$ROOT/include"
for f in tensor tables planner exec wisdom api_common dist dist_api; do gcc $CF -c $H/$f.c -o $OUT/$f.o & done
gcc $CF -c $H/api.c -o $OUT/api_d.o &
gcc $CF -DB2_SINGLE -c $H/api.c -o $OUT/api_f.o &
gcc $CF -c $H/f77api.c -o $OUT/f77_d.o &
gcc $CF -DB2_SINGLE -c $H/f77api.c -o $OUT/f77_f.o &
g++ -O2 -fPIC -std=c++17 -c $HERE/emu_shim.cpp -o $OUT/emu_shim.o &
wait
g++ -shared -o $OUT/libfftw3_b200_emu.so $OUT/*.o -lpthread -lm
