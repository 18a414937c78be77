#!/bin/bash
for c in 128 256 512 1024 2048 4096; do
  r=$(CB200_HOST_CHUNK=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['pcie_copy_only_ms'])")
  echo "chunk=$c e2e_ms value copy_only_ms: $r"
done
