#!/bin/bash
# Bounds check of the kernel sources: builds the host emulation of the kernels (csrc/emu.h, the
# harness behind tests/test_emu_kernels.py) with AddressSanitizer and runs the emulated-kernel
# tests under it.  Shared memory and "device" buffers are heap blocks in the emulation, so an
# out-of-range index in a kernel is reported with file:line.  (compute-sanitizer is not
# available on the GPU pool.)   usage: tools/asan_emu.sh [pytest args]
# SAN=thread tools/asan_emu.sh runs the same under ThreadSanitizer instead: every CUDA thread is an OS
# thread and __syncthreads a std::barrier there, so a missing barrier between a shared-memory write
# and a read by another thread is a reported data race (reports go to /tmp/chs_tsan.log.<pid>).
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/_build
[ -f tests/_build/libchs_emu.so ] && cp tests/_build/libchs_emu.so /tmp/libchs_emu.orig.so
SAN=${SAN:-address}
g++ -std=c++20 -O1 -g -fsanitize=$SAN -fno-omit-frame-pointer -ffp-contract=off -DCHS_EMU -x c++ \
    -shared -fPIC -pthread -o tests/_build/libchs_emu.so chsimpy_b200/csrc/chs_api.cu chsimpy_b200/csrc/chs_ll.cu
RT=$([ "$SAN" = thread ] && echo libtsan.so || echo libasan.so)
LD_PRELOAD=$(gcc -print-file-name=$RT) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
    TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4 log_path=/tmp/chs_tsan.log" \
    python -m pytest tests/test_emu_kernels.py tests/test_slab_ranks.py -x -q "$@" || rc=$?
[ -f /tmp/libchs_emu.orig.so ] && cp /tmp/libchs_emu.orig.so tests/_build/libchs_emu.so && touch tests/_build/libchs_emu.so
exit ${rc:-0}
