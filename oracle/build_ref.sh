#!/usr/bin/env bash
# Builds the UNMODIFIED reference kmer-db (v2.3.1) from the sources where they lie
# (REF, default /root/reference) into oracle/_ref/kmer-db.  Nothing is copied into
# the repo: objects and the binary go to oracle/_ref/ (git-ignored, shipped by gpurun).
# The reference's own make is not used (it needs network for submodules + nasm);
# we compile the few translation units directly against the system zlib.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  if [ -x "$OUT/kmer-db" ]; then echo "[build_ref] $REF absent; keeping prebuilt $OUT/kmer-db"; exit 0; fi
  echo "[build_ref] $REF absent and no prebuilt binary" >&2; exit 1
fi
mkdir -p "$OUT/obj" "$OUT/shim/zlib-ng"
# libs/refresh/compression/lib/file_wrapper.h includes <zlib-ng/zlib.h>; map it to system zlib
echo '#include <zlib.h>' > "$OUT/shim/zlib-ng/zlib.h"
FLAGS="-std=c++20 -O3 -mavx2 -m64 -DARCH_X64 -DREFRESH_USE_ZLIB -fpermissive -pthread -w -I$REF -I$REF/libs -I$OUT/shim"
SRCS=$(ls "$REF"/src/*.cpp "$REF"/src/kmc_api/*.cpp "$REF"/src/simd/*.cpp)
compile_one() { src="$1"; obj="$OUT/obj/$(echo "$src" | sed "s#$REF/##; s#/#_#g; s#\.cpp\$#.o#")";
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then g++ $FLAGS -c "$src" -o "$obj"; fi; }
export -f compile_one; export FLAGS OUT REF
echo "$SRCS" | xargs -P "$(nproc)" -I{} bash -c 'compile_one {}'
g++ -pthread "$OUT"/obj/*.o -lz -lpthread -o "$OUT/kmer-db"
echo "[build_ref] built $OUT/kmer-db"
