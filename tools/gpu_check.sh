#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), ncu launch list + full capture of the shooting kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh <tag>'
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.csv 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log
tail -3 gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$tag.log
tail -2 gpurun_out/smoke_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>&1
cat gpurun_out/bench_ref_$tag.json
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu1_$tag.log 2>&1
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shoot_kernel -s 8 -c 1 -f -o gpurun_out/prof_shoot_$tag $CMD > gpurun_out/ncu2_$tag.log 2>&1
ls -la gpurun_out
