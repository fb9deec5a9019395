#!/bin/bash
# Build library variants for A/B runs on one GPU box: tools/build_variants.sh name1="-DX=1 -DY=2" name2="..." ...
# Each variant is compiled from the current sources in its own object directory -> variants/<name>.so
# (variants/ is git-ignored but travels with gpurun snapshots). Run them with tools/ab.sh.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/julia-raytracer_b200/csrc
mkdir -p $ROOT/variants
NVFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall,-Wno-unused-function -diag-suppress 177 -ccbin /usr/bin/g++"
make -s -C $SRC   # host objects (jt_stage.o, jt_wide_bvh.o, jt_host_bvh.o) are shared by all variants
build_one() {
  name=${1%%=*}; flags=${1#*=}
  d=$(mktemp -d)
  for f in jt_api jt_group jt_probe jt_lights; do
    nvcc $NVFLAGS $flags -I$SRC -c $SRC/$f.cu -o $d/$f.o
  done
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/variants/$name.so $d/jt_api.o $d/jt_group.o $d/jt_probe.o $d/jt_lights.o \
     $SRC/jt_host_bvh.o $SRC/jt_wide_bvh.o $SRC/jt_stage.o $SRC/jt_host_scene.o -lz -cudart static -ccbin /usr/bin/g++ -Xcompiler -fopenmp,-pthread
  rm -rf $d
  echo "built variants/$name.so  [$flags]"
}
for v in "$@"; do build_one "$v" & 
  while [ $(jobs -r | wc -l) -ge 4 ]; do sleep 1; done
done
wait
