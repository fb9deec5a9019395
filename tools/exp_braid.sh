#!/bin/bash
# braid-size sweep on one box: Msamples/s + scene staging seconds per JT_BRAID_MAX value
mkdir -p gpurun_out
nproc
for sc in ecosys features1 classroom; do
for b in 0 64 16 1; do
  if [ "$sc" != ecosys ] && [ "$b" = 16 ]; then continue; fi
  JT_BRAID_MAX=$b timeout 600 python bench.py --scene $sc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/braid_${sc}_${b}.json 2>>gpurun_out/braid.err
  python - <<PY
import json
d=json.load(open("gpurun_out/braid_${sc}_${b}.json"))
print("$sc braid=$b", round(d["value"],1), "Msamples/s  e2e", round(d["e2e"]["value"],1), " upload_s", round(d["scene_upload"]["seconds"],2), " dev MB", d["scene_upload"]["device_bytes"]>>20, " mrays", round(d["mrays_per_s"]))
PY
done; done
