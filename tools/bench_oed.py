#!/usr/bin/env python
"""Device-resident timing of BASELINE config 3 (Lotka OED [x; G; F_triu], 64 intervals x 16384 parameter samples,
Fisher information + criteria):  python tools/bench_oed.py [--steps 20]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--intervals", type=int, default=64)
    ap.add_argument("--candidates", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import torch
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    spec = configs.config3_lotka_oed(a.intervals)
    lay, ps_t, st, N = configs.native_from_spec(spec, Tsit5(2))
    B, nvars = a.candidates, N.nvars
    rng = np.random.default_rng(configs.SEED0 + 3)
    ps = np.tile(to_vec(ps_t), (B, 1))
    blocks = N.block_structure()
    pp = rng.uniform(0.8, 1.2, (B, 2))
    for i in range(a.intervals):
        nu0 = 0 if i == 0 else 9
        ps[:, blocks[i] + nu0: blocks[i] + nu0 + 2] = pp
    dev = torch.device("cuda:0")
    sz = N.out_sizes()
    fields = ("fim", "criteria", "defects_kind")
    nsets = 3
    d_ps = [torch.from_numpy(np.ascontiguousarray(np.roll(ps, s, axis=0).T)).to(dev) for s in range(nsets)]
    outs = [{f: torch.empty((sz[f], B), dtype=torch.float64, device=dev) for f in fields} for _ in range(nsets)]
    fsum = torch.zeros(4, dtype=torch.float64, device=dev)
    N.set_stream(torch.cuda.current_stream().cuda_stream)
    want = nat.WANT_FIM | nat.WANT_CRIT | nat.WANT_DEFECTS
    flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA

    def step(s):
        k = s % nsets
        ptrs = {f: outs[k][f].data_ptr() for f in fields}
        ptrs["fim_sum"] = fsum.data_ptr()
        N.eval_raw(d_ps[k].data_ptr(), B, B, want, flags, ptrs, 1)

    for s in range(3):
        step(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(a.steps):
        step(s)
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / a.steps
    ms = float(np.mean(N.kernel_ms_history(a.steps)))
    units = a.intervals * B
    # SURVEY 8d worked example: E = 2 (12 + 1) = 26, C_f + C_JS + C_F = 51, nz = 9: W = 26*51 + 4*42*9 = 2838, Q = 288 B
    W, Q = 2838, 288
    peak = max(nat.probe_fp64_tflops(0, 20000), nat.probe_fp64_mma_tflops(0, 10000))
    print(json.dumps({"workload": "lotka_oed_64x16k", "intervals": a.intervals, "candidates": B, "kernel_ms": ms,
                      "ms_per_step": ms_step, "units_per_s": units / (ms_step * 1e-3), "flops_per_unit": W,
                      "achieved_tflops": W * units / (ms * 1e-3) / 1e12, "peak_tflops": peak,
                      "frac_fp64": W * units / (ms * 1e-3) / 1e12 / peak,
                      "achieved_gbs": Q * units / (ms * 1e-3) / 1e9, "frac_hbm": Q * units / (ms * 1e-3) / 1e9 / 6458.7}))


if __name__ == "__main__":
    main()
