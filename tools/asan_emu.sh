#!/bin/bash
# compute-sanitizer is closed on this GPU pool, so out-of-bounds / UB checking of the device code is done on
# the CPU: the device headers are compiled for the host (tests/emu) with AddressSanitizer + UBSan and driven
# through both traversal modes, both integrators and both samplers on six scenes.
set -e
cd "$(dirname "$0")/../tests/emu"
CSRC=../../julia-raytracer_b200/csrc
/usr/bin/g++ -O1 -g -std=c++17 -fPIC -fopenmp -ffp-contract=off -mfma -fsanitize=address,undefined \
  -fno-omit-frame-pointer -I/usr/local/cuda/include -include emu_shims.h -Wno-unused-function -Wno-attributes \
  -shared -o /tmp/libjt_emu_asan.so emu_device.cpp emu_error.cpp $CSRC/jt_stage.cpp $CSRC/jt_wide_bvh.cpp $CSRC/jt_host_bvh.cpp
cd ../..
LD_PRELOAD=$(/usr/bin/gcc -print-file-name=libasan.so):$(/usr/bin/gcc -print-file-name=libubsan.so) \
  ASAN_OPTIONS=detect_leaks=0 JT_EMU_LIB=/tmp/libjt_emu_asan.so python tools/asan_emu_case.py
