#!/usr/bin/env bash
# AddressSanitizer and ThreadSanitizer over the whole NUFFT translation unit (lightkurve_b200/csrc/ls_nufft.cu)
# running on the CUDA-on-CPU layer (tests/native/cuda_emu.h): exact-size workspaces, every thread of a block a host
# thread - out-of-bounds accesses and shared-memory races that a GPU would commit silently show up here.
# Round 1 result: both clean.   Usage: bash tools/nufft_sanitizers.sh
set -eu
cd "$(dirname "$0")/.."
out=$(mktemp -d)
for san in address thread; do
  g++ -std=c++17 -O1 -g -fsanitize=$san -fno-omit-frame-pointer -pthread -I/usr/local/cuda/include -Wno-attributes \
      -shared -fPIC -o "$out/libnufft_emu_$san.so" tests/native/nufft_emu_driver.cpp
  rt=$([ $san = address ] && gcc -print-file-name=libasan.so || gcc -print-file-name=libtsan.so)
  echo "=== $san sanitizer ==="
  LD_PRELOAD="$rt" ASAN_OPTIONS=detect_leaks=0 TSAN_OPTIONS="halt_on_error=0" \
      python tools/nufft_emulated_sanitizer_run.py "$out/libnufft_emu_$san.so" 2>&1 | grep -v '^$' | tail -15
done
