#!/bin/bash
# e2e of config 4 through host buffers against the chunk size of the host pipeline (default: 8 chunks = 8192 candidates)
for ch in 8192 4096 2048; do
  echo -n "CB200_HOST_CHUNK=$ch: "; CB200_HOST_CHUNK=$ch bash tools/gpu_quick.sh 4 --steps 3 | head -1 | cut -c1-130
done
