#!/bin/bash
# A/B of library variants on ONE box: tools/ab.sh "<scene> ..." <lib-or-'cur'> ...   (Msamples/s, kernel ms)
SCENES=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  for sc in $SCENES; do
    if [ "$v" = cur ]; then unset JTRACE_B200_LIB; else export JTRACE_B200_LIB=$PWD/variants/$v; fi
    timeout 300 python bench.py --scene $sc --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${v%.so}_${sc}.json 2>>gpurun_out/ab.err
    python - <<PY
import json
d=json.load(open("gpurun_out/ab_${v%.so}_${sc}.json"))
print("$v", "$sc", round(d["value"],1), "Msamples/s  e2e", round(d["e2e"]["value"],1), " launches", d["gpu_launches"], " ext_share", round(d["roofline"]["kernel_share_of_step"],3))
PY
  done
done
done
