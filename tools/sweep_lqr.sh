#!/bin/bash
# Sweep directions-per-thread and min-blocks builds of the LQR shooting kernel (bench config 2); OED with MINB 6.
out=gpurun_out/sweep_lqr2.txt; : > $out
for cfg in "libcorleone_b200.so 4 64" "libcb200_minb5.so 4 64" "libcb200_minb6.so 4 64" "libcb200_minb5.so 2 64" "libcb200_minb6.so 2 64" "libcb200_minb5.so 8 64"; do
    set -- $cfg
    r=$(CORLEONE_B200_LIB=$PWD/corleone.jl_b200/$1 CB200_DIRS_PER_THREAD=$2 CB200_BLOCK=$3 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac'])")
    echo "$1 G=$2 bs=$3 kernel_ms ms_per_step frac: $r" | tee -a $out
done
r=$(CORLEONE_B200_LIB=$PWD/corleone.jl_b200/libcb200_minb6.so python tools/bench_oed.py --steps 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], d['ms_per_step'], d['frac_fp64'])")
echo "OED minb6: $r" | tee -a $out
