#!/bin/bash
# One gpurun call (round 2): GPU tests, the bench line, ncu launch list of the bench command, --set full captures of the
# dominant kernel of configs 4, 3 and 2.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check_r2.sh <tag> [notest]'
tag=${1:-r2}
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log
  tail -8 gpurun_out/pytest_gpu_$tag.log
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$tag.log
  tail -2 gpurun_out/smoke_$tag.log
fi
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>&1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu1_$tag.log 2>&1
for cfg in "4:shoot_dense:--candidates 2960:1" "3:shoot_kernel:--steps 20:6" "2:shoot_kernel:--steps 20:6"; do
  IFS=: read only kre extra skip <<< "$cfg"
  C2="python bench.py --only $only --warmup 3 --no-cpu-baseline --steps 3 $extra"
  $C2 > gpurun_out/plain_${tag}_c$only.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/prof_${tag}_c$only $C2 > gpurun_out/ncu_${tag}_c$only.log 2>&1
done
ls -la gpurun_out | tail -20
