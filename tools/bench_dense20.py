#!/usr/bin/env python
"""Device-resident timing of BASELINE config 4 (dense 20-state problem, full sensitivity Jacobian) on a candidate
chunk:  python tools/bench_dense20.py [--intervals 256] [--candidates 2048] [--steps 5]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--intervals", type=int, default=256)
    ap.add_argument("--candidates", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--nx", type=int, default=20, help="states of the dense family: 16, 20, 24 or 32")
    a = ap.parse_args()
    import torch
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    spec = configs.config4_dense20(a.intervals, nx=a.nx)
    lay, ps_t, st, N = configs.native_from_spec(spec, Tsit5(2))
    B, nvars = a.candidates, N.nvars
    rng = np.random.default_rng(configs.SEED0 + 4)
    ps = np.tile(to_vec(ps_t), (B, 1)) + 0.1 * rng.standard_normal((B, nvars))
    dev = torch.device("cuda:0")
    d_ps = torch.from_numpy(np.ascontiguousarray(ps.T)).to(dev)
    sz = N.out_sizes()
    fields = ("sens", "defects_kind", "xend")
    outs = {f: torch.empty((sz[f], B), dtype=torch.float64, device=dev) for f in fields}
    N.set_stream(torch.cuda.current_stream().cuda_stream)
    want = nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_XEND
    flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA
    ptrs = {f: outs[f].data_ptr() for f in fields}
    for _ in range(2):
        N.eval_raw(d_ps.data_ptr(), B, B, want, flags, ptrs, 1)
    torch.cuda.synchronize()
    for _ in range(a.steps):
        N.eval_raw(d_ps.data_ptr(), B, B, want, flags, ptrs, 1)
    torch.cuda.synchronize()
    ms = float(np.mean(N.kernel_ms_history(a.steps)))
    units = a.intervals * B
    # E = M(6K+1) = 52, C_f = 2 nx^2 + 4 nx, C_JS = 2 nx^2 ns + 3 nx ns, nz = nx (1 + ns), ns = nx + 4
    # (nx = 20: C_f = 880, C_JS = 20 640, nz = 500; first interval: ns = 4)
    nx, ns = a.nx, a.nx + 4
    W = 52 * ((2 * nx * nx + 4 * nx) + (2 * nx * nx * ns + 3 * nx * ns)) + 8 * 42 * nx * (1 + ns)
    peak = max(nat.probe_fp64_tflops(0, 20000), nat.probe_fp64_mma_tflops(0, 10000))
    print(json.dumps({"workload": f"dense{a.nx}", "intervals": a.intervals, "candidates": B, "kernel_ms": ms,
                      "units_per_s": units / (ms * 1e-3), "flops_per_unit": W,
                      "achieved_tflops": W * units / (ms * 1e-3) / 1e12, "peak_tflops": peak,
                      "frac": W * units / (ms * 1e-3) / 1e12 / peak,
                      "sens_gbytes": 8 * sz["sens"] * B / 1e9}))


if __name__ == "__main__":
    main()
