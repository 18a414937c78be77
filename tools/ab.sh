#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 2"
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac'])"; }
echo "uni:   $($B 2>&1 | pr)"
echo "exact: $(CB200_EXACT_PIECE_STEPS=1 $B 2>&1 | pr)"
echo "uni:   $($B 2>&1 | pr)"
echo "exact: $(CB200_EXACT_PIECE_STEPS=1 $B 2>&1 | pr)"
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/plain_uni.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shoot_kernel -s 8 -c 1 -f -o gpurun_out/prof_shoot_uni $CMD > gpurun_out/ncu_uni.log 2>&1
