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
    # (4) the library's own data plane (cb200_comm_*): the bench's N > 1 workload in small -- dense 20-state model,
    #     candidates sharded over the ranks, allreduce(objective_sum, grad_sum) + allgather(defects) fused into the kernels
    #     (P2P transport) and through NCCL (fallback transport); compared with rank 0 evaluating the whole batch alone
    for transport in ("p2p", "nccl"):
        os.environ["CB200_COMM_TRANSPORT"] = transport
        spec = P.dense20_spec(8, np.random.default_rng(21))
        lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(2), device=local)
        Bl = 40
        B_total, b0 = Bl * world, Bl * rank
        rng2 = np.random.default_rng(22)
        ps = cb.to_vec(ps_t)[None, :] + 0.1 * rng2.standard_normal((B_total, N.nvars))
        ident = [nat.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        try:
            N.comm_init(world, rank, ident[0], max_B_total=B_total)
        except nat.NativeError as e:
            if transport == "p2p":
                print(f"rank {rank}: p2p transport unavailable here: {e}")
                ok = False
            raise
        info = N.comm_info()
        want = nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_XEND
        d_ps = torch.from_numpy(np.ascontiguousarray(ps[b0:b0 + Bl].T)).to(dev)
        d_sums = torch.zeros(1 + N.nvars, dtype=torch.float64, device=dev)
        d_gat = torch.full((N.n_defects, B_total), float("nan"), dtype=torch.float64, device=dev)
        d_x = torch.empty((N.nx * N.I, Bl), dtype=torch.float64, device=dev)
        N.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            N.eval_raw(d_ps.data_ptr(), Bl, Bl, want, nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA | nat.COMM_REDUCE | nat.COMM_GATHER,
                       dict(xend=d_x.data_ptr(), objective_sum=d_sums.data_ptr(), grad_sum=d_sums.data_ptr() + 8,
                            defects_gathered=d_gat.data_ptr()), 3, comm_B_total=B_total, comm_b0=b0)
        torch.cuda.synchronize()
        # host pointers as well (candidate-major), results in host memory
        rows = np.ascontiguousarray(ps[b0:b0 + Bl])
        h_gat, h_sums = np.full((B_total, N.n_defects), np.nan), np.zeros(1 + N.nvars)
        N.eval_raw(rows.ctypes.data, Bl, N.nvars, want, nat.COMM_REDUCE | nat.COMM_GATHER,
                   dict(objective_sum=h_sums.ctypes.data, grad_sum=h_sums.ctypes.data + 8, defects_gathered=h_gat.ctypes.data), 3,
                   comm_B_total=B_total, comm_b0=b0)
        sums = d_sums.cpu().numpy()
        same = [torch.zeros_like(d_sums) for _ in range(world)]
        dist.all_gather(same, d_sums)
        identical = all(bool(torch.equal(t, same[0])) for t in same)
        dist.barrier()
        N.comm_destroy()
        if rank == 0:
            lay1, ps_t1, st1, N1 = P.native_from_spec(spec, cb.Tsit5(2), device=local)
            full = N1.eval_host(ps, want, objective_row=3, with_sums=True)
            e1 = abs(sums[0] - full["objective_sum"][0]) / abs(full["objective_sum"][0])
            e2 = np.max(np.abs(sums[1:] - full["grad_sum"])) / max(1e-300, np.max(np.abs(full["grad_sum"])))
            e3 = float(np.max(np.abs(d_gat.cpu().numpy() - full["defects_kind"].T)))
            e4 = float(np.max(np.abs(h_gat - full["defects_kind"])))
            e5 = float(np.max(np.abs(h_sums - sums)))
            names = {1: "nccl", 2: "p2p"}
            print(f"comm dense20 B={B_total} world={world} transport={names.get(info['transport'])}: objective_sum rel {e1:.2e}, "
                  f"grad_sum rel {e2:.2e}, gathered defects max abs diff {e3:.1e} (device) {e4:.1e} (host), host sums vs device sums "
                  f"{e5:.1e}, sums bit-identical on all ranks: {identical}")
            ok &= e1 < 1e-12 and e2 < 1e-12 and e3 == 0.0 and e4 == 0.0 and e5 <= 1e-12 * np.max(np.abs(sums)) and identical
            ok &= names.get(info["transport"]) == transport
    os.environ.pop("CB200_COMM_TRANSPORT", None)
    # (5) the bench's N > 1 workload at FULL size (BASELINE config 5: 256 intervals x 4096 candidates of the dense 20-state
    #     model = 2^20 units, candidates sharded over the ranks, fused all-reduce + all-gather): the reduced sums and the
    #     gathered defect matrix against rank 0 evaluating the whole batch alone
    if os.environ.get("CB200_CHECK_FULL", "1") != "0" and 4096 % world == 0:
        from corleone_b200 import configs
        import bench
        spec = configs.config4_dense20(256)
        lay, ps_t, st, N = configs.native_from_spec(spec, cb.Tsit5(2), device=local)
        B_total = 4096
        Bl, b0 = B_total // world, rank * (B_total // world)
        full = bench.dense20_inputs_soa(torch, dev, cb.to_vec(ps_t), B_total, configs.SEED0 + 5)     # [nvars, B_total], same on every rank
        d_ps = full[:, b0:b0 + Bl].contiguous()
        ident = [nat.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        N.comm_init(world, rank, ident[0], max_B_total=B_total)
        want = nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_XEND | nat.WANT_OBJ | nat.WANT_GRAD
        flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA | nat.COMM_REDUCE | nat.COMM_GATHER
        d_sums = torch.zeros(1 + N.nvars, dtype=torch.float64, device=dev)
        d_x = torch.empty((N.nx * N.I, Bl), dtype=torch.float64, device=dev)
        d_obj = torch.empty((1, Bl), dtype=torch.float64, device=dev)
        N.set_stream(torch.cuda.current_stream().cuda_stream)
        n_def = N.n_defects
        gathered = torch.full((n_def, B_total), float("nan"), dtype=torch.float64, device=dev)
        for _ in range(2):
            N.eval_raw(d_ps.data_ptr(), Bl, Bl, want, flags,
                       dict(xend=d_x.data_ptr(), objective=d_obj.data_ptr(), objective_sum=d_sums.data_ptr(), grad_sum=d_sums.data_ptr() + 8,
                            defects_gathered=gathered.data_ptr()),
                       1, resident=nat.WANT_SENS, comm_B_total=B_total, comm_b0=b0)
        torch.cuda.synchronize()
        gp, gld = N.comm_gathered()
        assert gld == B_total and gp != 0
        same = [torch.zeros_like(d_sums) for _ in range(world)]
        dist.all_gather(same, d_sums)
        identical = all(bool(torch.equal(t, same[0])) for t in same)
        dist.barrier()
        N.comm_destroy()
        if rank == 0:
            lay1, ps_t1, st1, N1 = configs.native_from_spec(spec, cb.Tsit5(2), device=local)
            N1.set_stream(torch.cuda.current_stream().cuda_stream)
            r_dk = torch.empty((n_def, B_total), dtype=torch.float64, device=dev)
            r_sums = torch.zeros(1 + N1.nvars, dtype=torch.float64, device=dev)
            r_obj = torch.empty((1, B_total), dtype=torch.float64, device=dev)
            N1.eval_raw(full.data_ptr(), B_total, B_total, nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD, nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA,
                        dict(defects_kind=r_dk.data_ptr(), objective=r_obj.data_ptr(), objective_sum=r_sums.data_ptr(),
                             grad_sum=r_sums.data_ptr() + 8), 1)
            torch.cuda.synchronize()
            e_d = float((gathered - r_dk).abs().max().item())
            rs, ss = r_sums.cpu().numpy(), d_sums.cpu().numpy()
            e_o = abs(ss[0] - rs[0]) / abs(rs[0])
            e_g = np.max(np.abs(ss[1:] - rs[1:])) / np.max(np.abs(rs[1:]))
            e_l = float((d_obj - r_obj[:, b0:b0 + Bl]).abs().max().item())
            print(f"config 5 at full size (2^20 units) world={world}: gathered defects max abs diff {e_d:.1e}, objective_sum rel {e_o:.2e}, "
                  f"grad_sum rel {e_g:.2e}, local objectives max abs diff {e_l:.1e}, sums bit-identical on all ranks: {identical}")
            ok &= e_d == 0.0 and e_o < 1e-12 and e_g < 1e-12 and e_l == 0.0 and identical
        del full
    if rank == 0:
        print("multi_gpu_check", "ok" if ok else "FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
