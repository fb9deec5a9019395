for res in 640 1280 1810 2560; do for p in 1 2; do
JT_PIPELINES=$p timeout 300 python bench.py --resolution $res --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/exp1_${res}_${p}.json 2>gpurun_out/exp1.err
python - <<PY
import json;d=json.load(open("gpurun_out/exp1_${res}_${p}.json"));print("res",$res,"pipes",$p,round(d["value"],1),"Msamples/s e2e",round(d["e2e"]["value"],1),"ext share",round(d["roofline"]["kernel_share_of_step"],3),"launches",d["gpu_launches"])
PY
done; done
