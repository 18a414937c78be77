#!/bin/bash
# randomised parity sweep on one GPU: bash tools/gpu_fuzz.sh <first seed> <number of seeds> <cases per seed>
s0=${1:-100}; n=${2:-10}; c=${3:-100}
mkdir -p gpurun_out
for ((s=s0; s<s0+n; s++)); do
  timeout 900 python tests/fuzz_parity.py --cases $c --seed $s > gpurun_out/fuzz_$s.log 2>&1
  echo "seed $s rc=$? $(tail -1 gpurun_out/fuzz_$s.log | cut -c1-200)"; grep -c saveat gpurun_out/fuzz_$s.log; grep "BAD" gpurun_out/fuzz_$s.log | cut -c1-600 | head -3
done
