#!/bin/bash
mkdir -p gpurun_out
for sc in ecosys classroom features1 materials1 cornellbox; do
  JT_BUILD_VERBOSE=1 timeout 600 python bench.py --scene $sc --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/exp2_${sc}.json 2>gpurun_out/exp2_${sc}.err
  grep jt_build_wide gpurun_out/exp2_${sc}.err | tail -2
  python - <<PY
import json
d=json.load(open("gpurun_out/exp2_${sc}.json"))
print("$sc", round(d["value"],1), "Msamples/s  e2e", round(d["e2e"]["value"],1), " upload_s", round(d["scene_upload"]["seconds"],2), " dev MB", d["scene_upload"]["device_bytes"]>>20, " mrays", round(d["mrays_per_s"]))
PY
done
