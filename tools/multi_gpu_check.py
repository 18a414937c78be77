#!/usr/bin/env python
"""Sharded evaluation on real GPUs (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py

Every rank evaluates its contiguous candidate shard through cb200_eval on its own GPU; the cross-candidate sums
(objective, gradient, Fisher information) are combined with ONE all-reduce (corleone_b200.sharding); rank 0 then
evaluates the whole batch alone and compares.  The CPU-side logic of the same module is covered by the gloo tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import problems as P
    import corleone_b200 as cb
    from corleone_b200 import native as nat, sharding
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(7)                       # the same inputs on every rank
    ok = True
    # (1) NLP sums: LQR, objective + gradient summed over candidates
    spec = P.lqr_spec(20, rng)
    lay = P.product_layer(spec, cb.Tsit5(2), device=local)
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    B = 1001
    ps = cb.to_vec(ps_t)[None, :] + 0.1 * rng.standard_normal((B, N.nvars))
    want = nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_DEFECTS
    ev = sharding.ShardedEvaluator(lambda rows: N.eval_host(rows, want, objective_row=2, with_sums=True), device=dev)
    res = ev(ps, gather=("objective",))
    if rank == 0:
        full = N.eval_host(ps, want, objective_row=2, with_sums=True)
        e1 = abs(res["objective_sum"][0] - full["objective_sum"][0]) / abs(full["objective_sum"][0])
        e2 = np.max(np.abs(res["grad_sum"] - full["grad_sum"])) / np.max(np.abs(full["grad_sum"]))
        e3 = np.max(np.abs(res["objective"] - full["objective"]))
        print(f"lqr  B={B} world={world}: objective_sum rel {e1:.2e}, grad_sum rel {e2:.2e}, gathered objective max abs {e3:.2e}")
        ok &= e1 < 1e-12 and e2 < 1e-12 and e3 == 0.0
    # (2) Fisher information summed over parameter samples
    spec = P.config3_lotka_oed(16, rng)
    lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(2), device=local)
    B = 777
    ps = cb.to_vec(ps_t)[None, :] + 0.01 * rng.standard_normal((B, N.nvars))
    ev = sharding.ShardedEvaluator(lambda rows: N.eval_host(rows, nat.WANT_FIM), device=dev)
    res = ev(ps)
    if rank == 0:
        full = N.eval_host(ps, nat.WANT_FIM)
        e = np.max(np.abs(res["fim_sum"] - full["fim_sum"])) / np.max(np.abs(full["fim_sum"]))
        print(f"oed  B={B} world={world}: fim_sum rel {e:.2e}")
        ok &= e < 1e-12
    # (3) interval sharding for B < world size: one candidate, the intervals split across the GPUs, one all-reduce
    spec = P.lotka_ms24_spec(rng)
    lay = P.product_layer(spec, cb.Tsit5(4), device=local)
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    x = cb.to_vec(ps_t)[None, :] + 0.02 * rng.standard_normal((1, N.nvars))
    want = nat.WANT_XEND | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_JAC

    def evaluate_range(rows, first, last):
        N.set_interval_range(first, last)
        try:
            return N.eval_host(rows, want, objective_row=3)
        finally:
            N.set_interval_range(1, N.I)

    import time
    iev = sharding.IntervalShardedEvaluator(evaluate_range, N.I, device=dev)
    res = iev(x)
    t0 = time.perf_counter()
    for _ in range(20):
        iev(x)
    t_shard = (time.perf_counter() - t0) / 20
    if rank == 0:
        full = N.eval_host(x, want, objective_row=3)
        t0 = time.perf_counter()
        for _ in range(20):
            N.eval_host(x, want, objective_row=3)
        t_one = (time.perf_counter() - t0) / 20
        errs = {k: float(np.max(np.abs(res[k] - full[k]))) for k in ("xend", "defects_kind", "grad", "jac_vals")}
        eo = abs(res["objective"][0, 0] - full["objective"][0, 0]) / abs(full["objective"][0, 0])
        print(f"ms24 B=1 world={world} interval-sharded ({res['_shard'].size} intervals on rank 0): max abs diff {errs}, "
              f"objective rel {eo:.1e}; wall {t_shard * 1e6:.0f} us sharded vs {t_one * 1e6:.0f} us on one GPU")
        ok &= all(v == 0.0 for v in errs.values()) and eo < 1e-14
        print("multi_gpu_check", "ok" if ok else "FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
