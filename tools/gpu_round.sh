#!/bin/bash
# One GPU-box visit: build check, smoke, parity tests, bench (both arms + the other configs), ncu launch list + full capture.
# Usage (from the repo root on the box): bash tools/gpu_round.sh [tag] [quick]
TAG=${1:-r02}
QUICK=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu_${TAG}.csv 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "build+smoke exit $?" | tee -a gpurun_out/smoke_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_${TAG}.log
tail -5 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit $?"
cut -c1-400 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
[ -n "$QUICK" ] && exit 0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "bench ref exit $?"
cut -c1-300 gpurun_out/bench_ref_${TAG}.json
for sc in cornellbox features1 materials1 ecosys; do
  timeout 400 python bench.py --scene $sc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${sc}_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
  cut -c1-200 gpurun_out/bench_${sc}_${TAG}.json
done
timeout 300 python tools/bench_traverse.py classroom features1 ecosys > gpurun_out/traverse_${TAG}.json 2>>gpurun_out/bench_${TAG}.err
# profiler passes: only after the identical plain command has exited 0
PROF="python bench.py --steps 1 --warmup 1 --spp-per-step 4 --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "ncu launches exit $?"
# full capture in the steady state of a 512-spp step: skip the first ~2000 matching launches (ramp-up), take 3
PROF2="python bench.py --steps 1 --warmup 0 --spp-per-step 512 --no-cpu-baseline"
timeout 300 $PROF2 > gpurun_out/plain2_${TAG}.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_wf_(extend_persist|shade|probe)" -s 3000 -c 3 -o gpurun_out/prof_${TAG} -f $PROF2 > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full exit $?"
ls gpurun_out | head -60
