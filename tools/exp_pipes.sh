#!/bin/bash
# Pipelines x persistent-grid share sweep on one box (env knobs, no rebuild): tools/exp_pipes.sh "<scene> ..."
SCENES=${1:-classroom}
mkdir -p gpurun_out
for rep in 1 2; do
IFS=";" read -ra CFGLIST <<< "${CFGS:-2 1;1 1;3 1;4 1;3 2;4 2}"
for cfg in "${CFGLIST[@]}"; do
  set -- $cfg
  for sc in $SCENES; do
    JT_PIPELINES=$1 JT_PGRID_DIV=$2 timeout 300 python bench.py --scene $sc --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/pipes_$1_$2_$sc.json 2>>gpurun_out/pipes.err
    python - <<PY
import json
d=json.load(open("gpurun_out/pipes_$1_$2_$sc.json"))
print("pipelines $1 grid/$2", "$sc", round(d["value"],1), "Msamples/s  e2e", round(d["e2e"]["value"],1), " launches", d["gpu_launches"])
PY
  done
done
done
