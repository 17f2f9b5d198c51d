#!/bin/bash
# Bounds check of the kernel sources: builds the host emulation of the kernels (csrc/emu.h, the
# harness behind tests/test_emu_kernels.py) with AddressSanitizer and runs the emulated-kernel
# tests under it.  Shared memory and "device" buffers are heap blocks in the emulation, so an
# out-of-range index in a kernel is reported with file:line.  (compute-sanitizer is not
# available on the GPU pool.)   usage: tools/asan_emu.sh [pytest args]
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/_build
[ -f tests/_build/libchs_emu.so ] && cp tests/_build/libchs_emu.so /tmp/libchs_emu.orig.so
g++ -std=c++20 -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -DCHS_EMU -x c++ \
    -shared -fPIC -pthread -o tests/_build/libchs_emu.so chsimpy_b200/csrc/chs_api.cu
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
    python -m pytest tests/test_emu_kernels.py tests/test_slab_ranks.py -x -q "$@" || rc=$?
[ -f /tmp/libchs_emu.orig.so ] && cp /tmp/libchs_emu.orig.so tests/_build/libchs_emu.so && touch tests/_build/libchs_emu.so
exit ${rc:-0}
