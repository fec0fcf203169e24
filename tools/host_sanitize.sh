#!/usr/bin/env bash
# The host side (libkdbx_host.so and the CLI) built with -fsanitize=address,undefined and the CPU tests run against it.
# Needs a g++ that ships the sanitizer runtimes (/usr/bin/g++ in the build container; /opt/gcc does not) and an already
# built kmer-db_b200/lib/libkdbx.so.  The instrumented files replace the built ones for the run and are put back afterwards.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
PKG="$ROOT/kmer-db_b200"
CXX="${SAN_CXX:-/usr/bin/g++}"
ASAN="$("$CXX" -print-file-name=libasan.so)"; UBSAN="$("$CXX" -print-file-name=libubsan.so)"
SAVE="$(mktemp -d)"
cp "$PKG/lib/libkdbx_host.so" "$PKG/bin/kmer-db-b200" "$SAVE/"
restore() { cp "$SAVE/libkdbx_host.so" "$PKG/lib/"; cp "$SAVE/kmer-db-b200" "$PKG/bin/"; rm -rf "$SAVE"; }
trap restore EXIT
FLAGS="-O1 -g -std=c++17 -fPIC -Wall -Wextra -pthread -fsanitize=address,undefined -fno-omit-frame-pointer"
cd "$PKG"
"$CXX" $FLAGS -shared -o lib/libkdbx_host.so host/db_io.cpp host/csv_out.cpp host/synth.cpp host/host_api.cpp host/fasta.cpp \
  host/build.cpp host/params.cpp host/distance.cpp host/partition.cpp -Llib -lkdbx -lz -Wl,-rpath,'$ORIGIN'
"$CXX" $FLAGS -o bin/kmer-db-b200 host/main.cpp -Llib -lkdbx_host -lkdbx -Wl,-rpath,'$ORIGIN/../lib'
cd "$ROOT"
LD_PRELOAD="$(readlink -f "$ASAN") $(readlink -f "$UBSAN")" ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  python -m pytest tests/test_host.py tests/test_cli_host.py tests/test_multi_gpu_host.py tests/test_bench_contract.py -x -q -m "not gpu"
