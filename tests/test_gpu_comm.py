"""The multi-GPU data plane behind the C ABI (cb200_comm_*) -- exercised on ONE device: the ranks are handles of the same
layer in one process (peer windows by plain pointers), so the fused kernels (P2P defect stores of the shooting kernels,
one-shot all-reduce in the tail of the evaluation's last kernel) run exactly as they do across GPUs, minus NVLink.
tests/test_gpu_multi.py runs the same comparisons with one process per GPU (NCCL bootstrap, CUDA IPC) on >= 2 GPUs.

Reference counterpart: mythreadmap(::EnsembleDistributed) = pmap over intervals (src/Corleone.jl:24,
src/multiple_shooting.jl:118-130) followed by the merge (src/multiple_shooting.jl:139-182).
"""
import threading

import numpy as np
import pytest

import problems as P

pytestmark = pytest.mark.gpu


class ThreadGather:
    """the host's bootstrap all-gather between ranks that are threads of one process"""

    def __init__(self, n):
        self.n, self.slots, self.bar = n, [None] * n, threading.Barrier(n)

    def make(self, rank):
        def ag(data):
            self.slots[rank] = data
            self.bar.wait()
            out = list(self.slots)
            self.bar.wait()
            return out
        return ag


def make_ranks(spec, alg, nranks, max_B_total, device=0):
    import corleone_b200 as cb
    layers = []
    for r in range(nranks):
        lay = P.product_layer(spec, alg, device=device[r] if isinstance(device, (list, tuple)) else device)
        ps_t, st = cb.setup(None, lay)
        layers.append((lay, ps_t, lay._native(st, ps_t)))
    tg = ThreadGather(nranks)
    errs = [None] * nranks

    def init(r):
        try:
            layers[r][2].comm_init_host(nranks, r, tg.make(r), max_B_total)
        except Exception as e:   # noqa: BLE001
            errs[r] = e
            tg.bar.abort()

    ths = [threading.Thread(target=init, args=(r,)) for r in range(nranks)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not any(errs), errs
    return layers


@pytest.mark.parametrize("nranks,model", [(2, "lqr"), (3, "lqr"), (4, "dense20")])
def test_fused_allreduce_and_allgather_device_pointers(nranks, model):
    """N ranks on one device, unequal shards, device pointers, asynchronous launches from one host thread: the reduced
    sums and the gathered defects equal the single-handle evaluation of the whole batch"""
    import torch
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(5)
    spec = P.lqr_spec(12, rng) if model == "lqr" else P.dense20_spec(6, rng)
    alg = cb.Tsit5(2)
    obj_row = 2 if model == "lqr" else 3
    shards = [37, 28, 45, 19][:nranks]
    B_total = sum(shards)
    ranks = make_ranks(spec, alg, nranks, B_total)
    N0 = ranks[0][2]
    assert N0.comm_info()["transport"] == nat.TRANSPORT_P2P and N0.comm_info()["nranks"] == nranks
    ps = cb.to_vec(ranks[0][1])[None, :] + 0.05 * rng.standard_normal((B_total, N0.nvars))
    want = nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_JAC
    # single handle, whole batch (no communicator involved)
    lay1 = P.product_layer(spec, alg)
    ps_t1, st1 = cb.setup(None, lay1)
    N1 = lay1._native(st1, ps_t1)
    full = N1.eval_host(ps, want, objective_row=obj_row, with_sums=True)
    dev = torch.device("cuda:0")
    flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA | nat.COMM_REDUCE | nat.COMM_GATHER
    for rep in range(3):   # three epochs: the windows are double-buffered by epoch parity
        outs, b0 = [], 0
        streams = [torch.cuda.Stream() for _ in range(nranks)]
        # buffers first, launches second: a rank's tail spins until every peer has arrived, so nothing that synchronises
        # the device (allocation, pageable copies) may sit between the launches of one host thread
        for r, (lay, ps_t, N) in enumerate(ranks):
            B = shards[r]
            d_ps = torch.from_numpy(np.ascontiguousarray(ps[b0:b0 + B].T)).to(dev)
            d_obj = torch.empty((1, B), dtype=torch.float64, device=dev)
            d_jac = torch.empty((N.jac_nnz_var, B), dtype=torch.float64, device=dev)
            d_dk = torch.empty((N.n_defects, B), dtype=torch.float64, device=dev)
            d_sums = torch.full((1 + N.nvars,), np.nan, dtype=torch.float64, device=dev)
            d_gat = torch.full((N.n_defects, B_total), np.nan, dtype=torch.float64, device=dev)
            N.set_stream(streams[r].cuda_stream)
            if rep == 0:   # grow the handle's workspaces outside the collective sequence
                N.eval_raw(d_ps.data_ptr(), B, B, want, flags & ~(nat.COMM_REDUCE | nat.COMM_GATHER),
                           dict(objective=d_obj.data_ptr(), jac_vals=d_jac.data_ptr(), defects_kind=d_dk.data_ptr(),
                                objective_sum=d_sums.data_ptr(), grad_sum=d_sums.data_ptr() + 8), obj_row)
            outs.append((d_ps, d_obj, d_jac, d_dk, d_sums, d_gat, b0, B))
            b0 += B
        torch.cuda.synchronize()
        for r, (d_ps, d_obj, d_jac, d_dk, d_sums, d_gat, b0, B) in enumerate(outs):
            N = ranks[r][2]
            N.eval_raw(d_ps.data_ptr(), B, B, want, flags,
                       dict(objective=d_obj.data_ptr(), jac_vals=d_jac.data_ptr(), defects_kind=d_dk.data_ptr(),
                            objective_sum=d_sums.data_ptr(), grad_sum=d_sums.data_ptr() + 8, defects_gathered=d_gat.data_ptr()),
                       obj_row, comm_B_total=B_total, comm_b0=b0)
        torch.cuda.synchronize()
        ref_dk = full["defects_kind"].T          # [n_defects, B_total]
        sums0 = outs[0][4].cpu().numpy()
        for r, (d_ps, d_obj, d_jac, d_dk, d_sums, d_gat, b0, B) in enumerate(outs):
            N = ranks[r][2]
            assert N.comm_info()["error"] == 0
            s = d_sums.cpu().numpy()
            assert np.array_equal(s, sums0), "the reduced sums must be bit-identical on every rank (rank-ordered addition)"
            assert abs(s[0] - full["objective_sum"][0]) <= 1e-12 * abs(full["objective_sum"][0])
            assert np.max(np.abs(s[1:] - full["grad_sum"])) <= 1e-12 * max(1.0, np.max(np.abs(full["grad_sum"])))
            assert np.array_equal(d_gat.cpu().numpy(), ref_dk), f"gathered defects differ on rank {r}"
            assert np.array_equal(d_dk.cpu().numpy(), ref_dk[:, b0:b0 + B])          # the local view
            assert np.array_equal(d_obj.cpu().numpy()[0], full["objective"][b0:b0 + B, 0])
            assert np.array_equal(d_jac.cpu().numpy(), full["jac_vals"][b0:b0 + B].T)
            gp, gld = N.comm_gathered()
            assert gld == B_total and gp != 0
    for _, _, N in ranks:
        N.comm_destroy()
        N.close()


def test_comm_host_pointers_threads():
    """the same through host pointers (candidate-major, the Julia layout): every rank is a thread that calls cb200_eval and
    blocks until its results are back -- the all-reduced sums and the all-gathered defect matrix arrive in host memory"""
    import torch
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    # Needs one GPU per rank: a host-pointer evaluation blocks its thread on synchronous (pageable) copies and queues
    # device->host copies behind its tail, and ranks that SHARE a device share its copy engines and hardware queues
    # (CUDA_DEVICE_MAX_CONNECTIONS) -- a rank's copies can end up behind the peer's tail, which spins for this rank.
    # Here the ranks are threads of one process on two devices: windows by plain peer access (cudaDeviceEnablePeerAccess).
    if torch.cuda.device_count() < 2:
        pytest.skip("ranks with host pointers need a GPU each (tools/multi_gpu_check.py covers the same path with processes)")
    rng = np.random.default_rng(11)
    spec = P.lotka_spec(controls=rng.uniform(0, 1, 120), quadrature=True)
    alg = cb.Tsit5(2)
    nranks, shards = 2, [1500, 1203]          # the big shard takes the chunked, multi-stream host path
    B_total = sum(shards)
    ranks = make_ranks(spec, alg, nranks, B_total, device=[0, 1])
    N0 = ranks[0][2]
    ps = cb.to_vec(ranks[0][1])[None, :] + 0.05 * rng.standard_normal((B_total, N0.nvars))
    want = nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD
    lay1 = P.product_layer(spec, alg)
    ps_t1, st1 = cb.setup(None, lay1)
    full = lay1._native(st1, ps_t1).eval_host(ps, want, objective_row=3, with_sums=True)
    res = [None] * nranks
    ready = threading.Barrier(nranks)

    def work(r):
        N = ranks[r][2]
        b0 = sum(shards[:r])
        B = shards[r]
        rows = np.ascontiguousarray(ps[b0:b0 + B])
        out = dict(defects_kind=np.empty((B, N.n_defects)), objective=np.empty((B, 1)), objective_sum=np.zeros(1),
                   grad_sum=np.zeros(N.nvars), defects_gathered=np.full((B_total, N.n_defects), np.nan))
        # ranks that share ONE device (this test only): a first launch loads its kernel lazily, which synchronises the
        # device -- and would wait for a peer's tail that is spinning for this rank.  Load and allocate everything with
        # a non-collective call first; ranks on their own GPUs (tests/test_gpu_multi.py) need no such care
        # (with the WHOLE batch, so that no workspace has to grow -- free + allocate synchronises too -- later on)
        warm = dict(defects_kind=np.empty((B_total, N.n_defects)), objective=np.empty((B_total, 1)), objective_sum=np.zeros(1),
                    grad_sum=np.zeros(N.nvars))
        N.eval_raw(ps.ctypes.data, B_total, N.nvars, want, 0, {k: v.ctypes.data for k, v in warm.items()}, 3)
        ready.wait()
        for _ in range(2):
            N.eval_raw(rows.ctypes.data, B, N.nvars, want, nat.COMM_REDUCE | nat.COMM_GATHER,
                       {k: v.ctypes.data for k, v in out.items()}, 3, comm_B_total=B_total, comm_b0=b0)
        res[r] = out

    ths = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for r in range(nranks):
        b0 = sum(shards[:r])
        o = res[r]
        assert o is not None
        assert np.array_equal(o["defects_gathered"], full["defects_kind"])
        assert np.array_equal(o["defects_kind"], full["defects_kind"][b0:b0 + shards[r]])
        assert np.array_equal(o["objective_sum"], res[0]["objective_sum"]) and np.array_equal(o["grad_sum"], res[0]["grad_sum"])
        assert abs(o["objective_sum"][0] - full["objective_sum"][0]) <= 1e-12 * abs(full["objective_sum"][0])
        assert np.max(np.abs(o["grad_sum"] - full["grad_sum"])) <= 1e-12 * np.max(np.abs(full["grad_sum"]))
    for _, _, N in ranks:
        N.comm_destroy()


def test_fim_sum_allreduce_and_fused_readout():
    """config-3 shape: the Fisher information summed over the parameter samples of all ranks; the fused read-out of the
    shooting kernel agrees bit for bit with the stand-alone read-out kernel"""
    import os
    import torch
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(3)
    spec = P.config3_lotka_oed(8, rng)
    alg = cb.Tsit5(2)

    def build():
        return P.native_from_spec(spec, alg)

    lay, ps_t, st, N1 = build()
    B = 203
    ps = cb.to_vec(ps_t)[None, :] + 0.01 * rng.standard_normal((B, N1.nvars))
    want = nat.WANT_FIM | nat.WANT_CRIT
    fused = N1.eval_host(ps, want)
    os.environ["CB200_NO_FUSED_FIM"] = "1"
    try:
        plain = N1.eval_host(ps, want)
    finally:
        del os.environ["CB200_NO_FUSED_FIM"]
    assert np.array_equal(fused["fim"], plain["fim"]) and np.array_equal(fused["criteria"], plain["criteria"])
    assert np.max(np.abs(fused["fim_sum"] - plain["fim_sum"])) <= 1e-12 * np.max(np.abs(plain["fim_sum"]))
    # two ranks, FIM partial sums all-reduced
    nranks, shards = 2, [120, 83]
    ranks = []
    for _ in range(nranks):
        ranks.append(build())
    tg = ThreadGather(nranks)
    ths = [threading.Thread(target=lambda r=r: ranks[r][3].comm_init_host(nranks, r, tg.make(r), 0)) for r in range(nranks)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    dev = torch.device("cuda:0")
    outs, b0 = [], 0
    fl = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA
    for r in range(nranks):   # buffers and a non-collective warm-up first (nothing that synchronises the device between the launches)
        N = ranks[r][3]
        Br = shards[r]
        d_ps = torch.from_numpy(np.ascontiguousarray(ps[b0:b0 + Br].T)).to(dev)
        d_fim = torch.empty((4, Br), dtype=torch.float64, device=dev)
        d_fs = torch.full((4,), np.nan, dtype=torch.float64, device=dev)
        N.eval_raw(d_ps.data_ptr(), Br, Br, nat.WANT_FIM, fl, dict(fim=d_fim.data_ptr(), fim_sum=d_fs.data_ptr()), 1)
        outs.append((d_ps, d_fim, d_fs))
        b0 += Br
    torch.cuda.synchronize()
    for r in range(nranks):
        d_ps, d_fim, d_fs = outs[r]
        ranks[r][3].eval_raw(d_ps.data_ptr(), shards[r], shards[r], nat.WANT_FIM, fl | nat.COMM_REDUCE,
                             dict(fim=d_fim.data_ptr(), fim_sum=d_fs.data_ptr()), 1)
    torch.cuda.synchronize()
    for r in range(nranks):
        s = outs[r][2].cpu().numpy()
        assert np.array_equal(s, outs[0][2].cpu().numpy())
        assert np.max(np.abs(s - fused["fim_sum"])) <= 1e-12 * np.max(np.abs(fused["fim_sum"]))
        ranks[r][3].comm_destroy()


def test_collective_flags_need_a_communicator():
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    spec = P.lqr_spec(4)
    lay = P.product_layer(spec, cb.Tsit5(2))
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    x = cb.to_vec(ps_t)[None, :].copy()
    out = np.zeros(1)
    with pytest.raises(nat.NativeError, match="communicator"):
        N.eval_raw(x.ctypes.data, 1, N.nvars, nat.WANT_OBJ, nat.COMM_REDUCE, dict(objective_sum=out.ctypes.data), 2)
    assert N.comm_info() == dict(nranks=1, rank=0, transport=0, epoch=0, max_B_total=0, error=0)


def test_resident_outputs_stay_on_the_device():
    """cb200_out.resident: the constraint-Jacobian values and sensitivities of a host-pointer evaluation stay in
    handle-owned device memory, quantity-major [q * B + b]; the host receives only what it asked for"""
    import torch
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(2)
    spec = P.lqr_spec(10, rng)
    lay = P.product_layer(spec, cb.Tsit5(2))
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    for B in (3, 1300):          # the small pinned path and the chunked multi-stream path
        ps = cb.to_vec(ps_t)[None, :] + 0.05 * rng.standard_normal((B, N.nvars))
        want = nat.WANT_DEFECTS | nat.WANT_JAC | nat.WANT_SENS | nat.WANT_OBJ
        full = N.eval_host(ps, want, objective_row=2)
        dk, obj = np.empty((B, N.n_defects)), np.empty((B, 1))
        N.eval_raw(ps.ctypes.data, B, N.nvars, want, 0, dict(defects_kind=dk.ctypes.data, objective=obj.ctypes.data), 2,
                   resident=nat.WANT_JAC | nat.WANT_SENS)
        assert np.array_equal(dk, full["defects_kind"]) and np.array_equal(obj, full["objective"])
        for bit, key, n in ((nat.WANT_JAC, "jac_vals", N.jac_nnz_var), (nat.WANT_SENS, "sens", N.n_sens)):
            ptr, ld = N.resident(bit)
            assert ld == B and ptr != 0
            t = torch.empty((n, B), dtype=torch.float64, device="cuda:0")
            import ctypes
            cudart = ctypes.CDLL("libcudart.so")
            assert cudart.cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(8 * n * B), 3) == 0
            assert np.array_equal(t.cpu().numpy().T, full[key])
        with pytest.raises(nat.NativeError):
            N.resident(nat.WANT_GRAD)          # not produced by the last evaluation


def test_small_evaluation_graph_replay_matches_and_survives_reallocation():
    """B = 1 through host pointers is one captured CUDA graph (config 1's latency path): replays give the same numbers,
    different inputs give different numbers, and a larger evaluation in between (workspaces move) does not break it"""
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(4)
    spec = P.lotka_ms24_spec(rng)
    lay = P.product_layer(spec, cb.Tsit5(4))
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    want = nat.WANT_JAC | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD
    x1 = cb.to_vec(ps_t)[None, :] + 0.02 * rng.standard_normal((1, N.nvars))
    x2 = cb.to_vec(ps_t)[None, :] + 0.02 * rng.standard_normal((1, N.nvars))
    import os
    os.environ["CB200_NO_GRAPH"] = "1"
    try:
        lay0 = P.product_layer(spec, cb.Tsit5(4))
        ps_t0, st0 = cb.setup(None, lay0)
        N0 = lay0._native(st0, ps_t0)          # a handle that never captures
    finally:
        del os.environ["CB200_NO_GRAPH"]
    ref1, ref2 = N0.eval_host(x1, want, objective_row=3), N0.eval_host(x2, want, objective_row=3)
    keys = ("jac_vals", "defects_kind", "objective", "grad")
    l0 = N.launch_count()
    a = N.eval_host(x1, want, objective_row=3)          # captures
    per_call = N.launch_count() - l0
    b = N.eval_host(x2, want, objective_row=3)          # replays
    c = N.eval_host(x1, want, objective_row=3)
    assert N.launch_count() - l0 == 3 * per_call          # the launch counter keeps counting the kernels of a replay
    big = cb.to_vec(ps_t)[None, :] + 0.02 * rng.standard_normal((700, N.nvars))
    N.eval_host(big, want | nat.WANT_SENS | nat.WANT_XEND, objective_row=3)          # grows workspaces: graphs are dropped
    d = N.eval_host(x2, want, objective_row=3)
    for k in keys:
        assert np.array_equal(a[k], ref1[k]) and np.array_equal(c[k], ref1[k]), k
        assert np.array_equal(b[k], ref2[k]) and np.array_equal(d[k], ref2[k]), k
    assert not np.array_equal(ref1["objective"], ref2["objective"])


@pytest.mark.parametrize("model,G", [("lqr", 8), ("lotka3", 5)])
def test_wide_direction_groups_run_every_live_direction(model, G):
    """with G > 4 directions per thread the prefix-length chain steps by two; a piece must run the smallest
    instantiated length >= its live prefix (na = 7 of G = 8, na = 4 of G = 5 ...), not the largest one below it"""
    import os
    import test_gpu_parity as tp
    import corleone_b200 as cb
    rng = np.random.default_rng(9)
    os.environ["CB200_DIRS_PER_THREAD"] = str(G)
    try:
        if model == "lqr":
            # 7 control pieces per interval: with (x, a, b) leading, the live prefix grows 4, 5, 6, 7, 8 across the pieces
            n = 14
            t = np.arange(n) / 10.0
            spec = dict(model="lqr", u0=[1.0, 0.0], p0=[-1.0, 1.0, 0.0], tspan=(0.0, n / 10.0),
                        controls=[(3, t, rng.standard_normal(n))], tunable_ic=[], quadrature_indices=[],
                        shooting_points=[0.0, 0.7])
            errs = tp.full_compare(spec, cb.Tsit5(2), 37, rng, objective_row=2)
        else:
            spec = P.lotka_spec(shooting_points=(0.0, 0.4, 0.8), controls=rng.uniform(0, 1, 120), tunable_ic=(1,))
            spec["tspan"] = (0.0, 1.2)
            errs = tp.full_compare(spec, cb.Tsit5(2), 33, rng, objective_row=3)
    finally:
        del os.environ["CB200_DIRS_PER_THREAD"]
    tp.assert_ok(errs)
