#!/usr/bin/env python
"""Device-resident timing of the adaptive mode (cb200_method 2) next to the fixed-step mode on the same layer:
BASELINE config 1's Lotka layer (24 intervals, 5 pieces each) for a batch of candidates.

    python tools/bench_adaptive.py [--candidates 4096]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    import torch
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    rng = np.random.default_rng(1)
    out = {}
    for name, alg, kw in (("fixed K=4", Tsit5(4), {}), ("adaptive 1e-8/1e-6", Tsit5(adaptive=True), dict(abstol=1e-8, reltol=1e-6)),
                          ("adaptive 1e-6/1e-3", Tsit5(adaptive=True), {})):
        spec = dict(configs.config1_lotka_ms24(np.random.default_rng(1)), kwargs=kw)
        lay, ps_t, st, N = configs.native_from_spec(spec, alg)
        B = a.candidates
        ps = to_vec(ps_t)[None, :] + 0.02 * rng.standard_normal((B, N.nvars))
        d_ps = torch.from_numpy(np.ascontiguousarray(ps.T)).cuda()
        sz = N.out_sizes()
        fields = ("defects_kind", "objective", "grad", "jac_vals")
        outs = {f: torch.empty((sz[f], B), dtype=torch.float64, device="cuda") for f in fields}
        N.set_stream(torch.cuda.current_stream().cuda_stream)
        want = nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_JAC
        flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA
        ptrs = {f: outs[f].data_ptr() for f in fields}
        for _ in range(3):
            N.eval_raw(d_ps.data_ptr(), B, B, want, flags, ptrs, 3)
        torch.cuda.synchronize()
        for _ in range(a.steps):
            N.eval_raw(d_ps.data_ptr(), B, B, want, flags, ptrs, 3)
        torch.cuda.synchronize()
        ms = float(np.mean(N.kernel_ms_history(a.steps)))
        out[name] = dict(kernel_ms=ms, units_per_s=N.I * B / ms * 1e3)
    print(json.dumps(dict(workload="lotka_ms24 layer, f + c + grad + cons_j", candidates=a.candidates, **out)))


if __name__ == "__main__":
    main()
