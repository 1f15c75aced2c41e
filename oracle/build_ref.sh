#!/bin/sh
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compile the UNMODIFIED reference sources (in place, from $REF) into
# oracle/_ref/libfftw3{,f,l}_ref.so with the codelet tables of refbuild/codelets_gen.c (see oracle/Makefile header).
# usage: build_ref.sh <suffix: "" | f | l>
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
SFX=$1
case "$SFX" in
  "") PREC="" ;;
  f)  PREC="-DFFTW_SINGLE=1" ;;
  l)  PREC="-DFFTW_LDOUBLE=1" ;;
  *)  echo "bad suffix"; exit 2 ;;
esac
OBJ=$OUT/obj/x$SFX
mkdir -p "$OBJ"
CFLAGS="-O3 -march=x86-64-v3 -fPIC -fopenmp -w -I$HERE/refbuild -I$REF -I$REF/kernel -I$REF/dft -I$REF/rdft -I$REF/reodft -I$REF/api -I$REF/threads -I$REF/dft/scalar -I$REF/rdft/scalar -I$REF/simd-support $PREC"
SRCS=""
for d in kernel dft dft/scalar rdft rdft/scalar reodft api; do
  SRCS="$SRCS $(ls $REF/$d/*.c)"
done
for f in api.c conf.c ct.c dft-vrank-geq1.c hc2hc.c rdft-vrank-geq1.c vrank-geq1-rdft2.c openmp.c; do
  SRCS="$SRCS $REF/threads/$f"
done
SRCS="$SRCS $HERE/refbuild/codelets_gen.c"
for f in $SRCS; do
  o=$OBJ/$(echo "$f" | sed "s#^$REF/##; s#^$HERE/##; s#/#_#g; s#\.c\$#.o#")
  echo "gcc $CFLAGS -c $f -o $o"
done | xargs -P "$(nproc)" -I{} sh -c '{}'
gcc -shared -fopenmp -o "$OUT/libfftw3${SFX}_ref.so" "$OBJ"/*.o -lm
echo "built $OUT/libfftw3${SFX}_ref.so"
