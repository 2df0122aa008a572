#!/bin/sh
# TEST INFRASTRUCTURE ONLY.  Builds the *unmodified* reference (AntiZ main.cpp +
# its vendored zlib 1.2.8) from the sources where they lie under /root/reference
# into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
# No reference source is copied into this repository.  Recipe = SURVEY.md App. C.
#   oracle/_ref/libz128.so   - zlib 1.2.8 (deflate/inflate/adler32), for ctypes
#   oracle/_ref/uncomp_ref   - the reference `uncomp` binary
# If /root/reference is absent (GPU box) this script is a no-op: the prebuilt
# files that came with the snapshot are used.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
R="${ANTIZ_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$R" ]; then
  echo "build_ref: $R not present; keeping prebuilt $OUT" ; exit 0
fi
Z="$R/includes, tools, stuff/zlib test/zlib128"
T="$R/includes, tools, stuff/tclap/tclap-1.2.1/include"
mkdir -p "$OUT/zobj" "$OUT/shim"
ln -sf "$R/ATZData.h" "$OUT/shim/AtzData.h"     # main.cpp:1 includes "AtzData.h"
OBJS=""
for f in adler32 compress crc32 deflate infback inffast inflate inftrees trees uncompr zutil; do
  gcc -O3 -fPIC -w -c "$Z/$f.c" -o "$OUT/zobj/$f.o"
  OBJS="$OBJS $OUT/zobj/$f.o"
done
ar rcs "$OUT/libz128.a" $OBJS
gcc -shared -Wl,-Bsymbolic -o "$OUT/libz128.so" $OBJS   # -Bsymbolic: python already maps the system libz 1.3
g++ -std=c++14 -O3 -w -DHAVE_LONG_LONG -include cstring -I "$OUT/shim" -I "$R" -I "$Z" -I "$T" \
    "$R/main.cpp" "$OUT/libz128.a" -o "$OUT/uncomp_ref"
echo "build_ref: built $OUT/libz128.so and $OUT/uncomp_ref"
