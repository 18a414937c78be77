#!/bin/bash
# Sweep directions-per-thread, block size and min-blocks builds of the LQR shooting kernel (bench config 2).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
out=gpurun_out/sweep_lqr.txt; : > $out
for cfg in "libcorleone_b200.so 8 64" "libcorleone_b200.so 4 64" "libcb200_minb3.so 8 64" "libcb200_minb3.so 4 64" "libcb200_minb3.so 4 128"  "libcb200_minb4.so 4 64" "libcb200_minb4.so 4 128" "libcb200_minb4.so 4 32" "libcb200_minb4.so 2 128" "libcb200_minb4.so 2 64"; do
    set -- $cfg
    r=$(CORLEONE_B200_LIB=$PWD/corleone.jl_b200/$1 CB200_DIRS_PER_THREAD=$2 CB200_BLOCK=$3 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])")
    echo "$1 G=$2 bs=$3 kernel_ms ms_per_step frac launches: $r" | tee -a $out
done
