#!/bin/bash
# what the fused tail costs per evaluation: config-2 weak scaling (0.12 ms evaluations) on N GPUs with the full tail, flags only,
# and a local tail.  gpurun --gpus 2 -- 'bash tools/gpu_tail_probe.sh 2'
n=${1:-2}
for probe in 0 1 2; do
  CB200_TAIL_PROBE=$probe timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+probe)) bench.py --gpus $n --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['secondary']
print('probe $probe: lqr weak ms/step %.4f  value %.4g  e2e sums only %.4g' % (s['ms_per_step'], s['value'], s['e2e_sums_only']['value']))"
done
