#!/usr/bin/env python
"""Where does config 4's e2e time go?  cb200_eval with pinned host buffers for different output sets (the sensitivity Jacobian
always stays resident): python tools/e2e_probe.py [--candidates 16384]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=16384)
    a = ap.parse_args()
    import torch
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    spec = configs.config4_dense20(256)
    lay, ps_t, st, N = configs.native_from_spec(spec, Tsit5(2))
    B, nvars = a.candidates, N.nvars
    sz = N.out_sizes()
    rng = np.random.default_rng(1)
    pin_ps = torch.empty((B, nvars), dtype=torch.float64, pin_memory=True)
    pin_ps.copy_(torch.from_numpy(np.tile(to_vec(ps_t), (B, 1)) + 0.1 * rng.standard_normal((B, nvars))))
    pin = {f: torch.empty((B, sz[f]), dtype=torch.float64, pin_memory=True) for f in ("defects_kind", "xend")}
    d_ps = pin_ps.t().contiguous().cuda()
    N.set_stream(torch.cuda.current_stream().cuda_stream)
    flags_dev = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA

    def timeit(fn, n=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    res = nat.WANT_SENS
    t_dev = timeit(lambda: N.eval_raw(d_ps.data_ptr(), B, B, nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_XEND, flags_dev, {}, 1,
                                      resident=nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_XEND))
    print(f"device-resident, everything stays:           {t_dev:8.2f} ms")
    for name, want, fields in (("host in, nothing out (all resident)", nat.WANT_SENS, ()),
                               ("host in, defects out", nat.WANT_SENS | nat.WANT_DEFECTS, ("defects_kind",)),
                               ("host in, defects + xend out (the bench)", nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_XEND, ("defects_kind", "xend"))):
        t = timeit(lambda: N.eval_raw(pin_ps.data_ptr(), B, nvars, want, 0, {f: pin[f].data_ptr() for f in fields}, 1, resident=res))
        print(f"{name:45s}{t:8.2f} ms   (+{t - t_dev:6.2f})")
    for ch in (1024, 4096, 16384):
        os.environ["CB200_HOST_CHUNK"] = str(ch)
        t = timeit(lambda: N.eval_raw(pin_ps.data_ptr(), B, nvars, nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_XEND, 0,
                                      {f: pin[f].data_ptr() for f in ("defects_kind", "xend")}, 1, resident=res))
        print(f"bench outputs, CB200_HOST_CHUNK={ch:<6d}         {t:8.2f} ms   (+{t - t_dev:6.2f})")


if __name__ == "__main__":
    main()
