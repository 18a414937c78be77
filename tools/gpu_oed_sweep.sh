#!/bin/bash
for mb in 5 4 6; do for bs in 64 128 32; do
  echo -n "MINB=$mb BLOCK=$bs: "; CB200_OED_MINB=$mb CB200_BLOCK=$bs bash tools/gpu_quick.sh 3 | cut -c1-110
done; done
