#!/bin/bash
# Strong-scaling sweep on one multi-GPU box: gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale.sh <tag> "2 4 8"'
# multi_gpu_check (all collectives against the single-GPU results, incl. config 5 at full size) on the largest N first,
# then bench.py --gpus N for every N.
tag=${1:-r2}; ns=${2:-"2 4 8"}
mkdir -p gpurun_out
nmax=$(echo $ns | awk '{print $NF}')
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $nmax --master-addr 127.0.0.1 --master-port 29541 tools/multi_gpu_check.py > gpurun_out/multi_gpu_check_${tag}_n$nmax.log 2>&1
echo "multi_gpu_check rc=$?"; grep -E "world=|multi_gpu_check" gpurun_out/multi_gpu_check_${tag}_n$nmax.log | cut -c1-400
port=29600
for n in $ns; do
  port=$((port+1))
  extra="--no-secondary"; [ "$n" = "$nmax" ] && extra=""
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 $extra > gpurun_out/bench_${tag}_n$n.json 2> gpurun_out/bench_${tag}_n$n.err
  echo "bench n=$n rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_n$n.json").read().strip().splitlines()[-1])
    print("N=%d value %.4g units/s  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.4g  transport %s  clocks %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["config_details"]["transport"][:8], d["clocks"]["sm_mhz"]))
    if d.get("secondary"): print("   secondary", d["secondary"]["workload"][:40], "%.4g" % d["secondary"]["value"], "e2e sums only %.4g" % d["secondary"]["e2e_sums_only"]["value"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/bench_${tag}_n$n.err").read()[-1500:])
PY
done
