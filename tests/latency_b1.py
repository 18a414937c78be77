#!/usr/bin/env python
"""Single-candidate latency of cb200_eval (BASELINE config 1: Lotka fishing, 24 intervals, Tsit5 K=4, 10 sensitivity
directions): f, grad f, c, cons_j through host pointers, as one NLP iterate of the reference would ask for."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # lives with the tests: times the CPU checker too


def main():
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    from oracle import pyoracle as po
    spec = configs.config1_lotka_ms24()
    lay, ps_t, st, N = configs.native_from_spec(spec, Tsit5(4))
    ps = to_vec(ps_t)[None, :].copy()
    sz = N.out_sizes()
    fields = ("jac_vals", "defects_kind", "objective", "grad")
    out = {f: np.empty((1, sz[f])) for f in fields}
    ptrs = {f: out[f].ctypes.data for f in fields}
    want = nat.WANT_JAC | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD
    for _ in range(50):
        N.eval_raw(ps.ctypes.data, 1, N.nvars, want, 0, ptrs, 3)
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        N.eval_raw(ps.ctypes.data, 1, N.nvars, want, 0, ptrs, 3)
    gpu_us = (time.perf_counter() - t0) / n * 1e6
    O = po.OracleLayer(spec)
    t0 = time.perf_counter()
    m = 200
    for _ in range(m):
        O.bench_units(ps, K=4, with_sens=True, nthreads=1)
    cpu_us = (time.perf_counter() - t0) / m * 1e6
    print(json.dumps({"workload": "lotka_ms24 (config 1), B = 1", "nvars": N.nvars, "cb200_eval_us": gpu_us,
                      "kernel_us": 1e3 * N.last_kernel_ms(), "oracle_1thread_us": cpu_us}))


if __name__ == "__main__":
    main()
