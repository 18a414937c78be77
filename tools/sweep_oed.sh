#!/bin/bash
for lib in libcorleone_b200.so libcb200_minb3.so libcb200_minb4.so libcb200_minb5.so; do
 for bs in 64 128; do
  r=$(CORLEONE_B200_LIB=$PWD/corleone.jl_b200/$lib CB200_BLOCK=$bs python tools/bench_oed.py --steps 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], d['ms_per_step'], d['frac_fp64'])")
  echo "$lib bs=$bs kernel_ms ms_per_step frac: $r"
 done
done
