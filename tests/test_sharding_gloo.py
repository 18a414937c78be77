"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous candidate shards, one all-reduce of
the cross-candidate sums, all-gather of per-candidate rows.  The rank-local evaluation is played by the CPU
oracle here (the GPU path is covered by tests/test_gpu_parity.py); what is under test is that sharded
evaluation + collectives reproduce the unsharded result."""
import os
import socket

import numpy as np
import pytest

import problems as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_eval_factory():
    from oracle import pyoracle as po
    rng = np.random.default_rng(5)
    spec = P.lotka_spec(controls=rng.uniform(0, 1, 120), quadrature=True)
    O = po.OracleLayer(spec)

    def evaluate_local(ps_rows):
        obj, grad, defk = [], [], []
        for row in ps_rows:
            o = O.eval(row, K=1)
            _, _, Jl = O.jacobians(row, K=1)
            obj.append([o["last"][2]])
            grad.append(Jl[2])
            defk.append(o["defk"])
        obj, grad = np.array(obj).reshape(-1, 1), np.array(grad).reshape(-1, O.nvars)
        return dict(objective=obj, grad=grad, defects_kind=np.array(defk).reshape(-1, O.ndef),
                    objective_sum=obj.sum(axis=0), grad_sum=grad.sum(axis=0))

    return O, evaluate_local


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    from corleone_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O, ev = _oracle_eval_factory()
        rng = np.random.default_rng(99)
        ps = O.default_ps()[None, :] + 0.05 * rng.standard_normal((B, O.nvars))
        res = sharding.ShardedEvaluator(ev)(ps, gather=("objective", "defects_kind"))
        q.put((rank, res["_shard"].start, res["_shard"].stop, res["objective_sum"], res["grad_sum"],
               res["objective"], res["defects_kind"], res["grad"].shape[0]))
    finally:
        dist.destroy_process_group()


def test_partition_properties():
    from corleone_b200.sharding import counts, partition
    for B in (0, 1, 5, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            sh = [partition(B, world, r) for r in range(world)]
            assert sh[0].start == 0 and sh[-1].stop == B
            assert all(a.stop == b.start for a, b in zip(sh[:-1], sh[1:]))
            assert max(s.size for s in sh) - min(s.size for s in sh) <= 1
            assert counts(B, world) == [s.size for s in sh]
    with pytest.raises(ValueError):
        partition(4, 2, 2)


@pytest.mark.parametrize("B", [5, 8])
def test_world2_gloo_matches_unsharded(B):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    O, ev = _oracle_eval_factory()
    rng = np.random.default_rng(99)
    ps = O.default_ps()[None, :] + 0.05 * rng.standard_normal((B, O.nvars))
    full = ev(ps)
    out.sort()
    assert [(o[1], o[2]) for o in out] == [(0, (B + 1) // 2), ((B + 1) // 2, B)]
    for o in out:
        assert np.allclose(o[3], full["objective_sum"], rtol=1e-13, atol=0)
        assert np.allclose(o[4], full["grad_sum"], rtol=1e-12, atol=1e-13)
        assert np.array_equal(o[5], full["objective"])          # gathered rows are exact copies, global order
        assert np.array_equal(o[6], full["defects_kind"])
        assert o[7] == o[2] - o[1]                               # per-candidate gradient stays rank-local


def _interval_worker(rank, world, port, B, q):
    """interval sharding: the rank-local evaluation is the oracle masked to the rank's interval range"""
    import torch.distributed as dist
    from corleone_b200 import sharding
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        spec = P.lotka_spec(shooting_points=[0.0, 2.0, 4.0, 6.0, 8.0, 10.0], controls=rng.uniform(0, 1, 120), quadrature=True)
        O = po.OracleLayer(spec)
        blocks = O.block_structure()
        per = O.ndef // (O.I - 1)        # kind-major u0 matchings only (no tunable p, no control matchings here)

        def evaluate_range(ps_rows, first, last):
            own = np.zeros(O.I, dtype=bool)
            own[first - 1:last] = True
            xend = np.zeros((len(ps_rows), O.nx * O.I))
            defk = np.zeros((len(ps_rows), O.ndef))
            obj = np.zeros((len(ps_rows), 1))
            grad = np.zeros((len(ps_rows), O.nvars))
            for b, row in enumerate(ps_rows):
                o = O.eval(row, K=1)
                Jx, _, _ = O.jacobians(row, K=1)
                for i in np.nonzero(own)[0]:
                    xend[b, O.nx * i:O.nx * (i + 1)] = o["xend"][:, i]
                    if i + 1 < O.I:   # the defect at node i+1 belongs to the interval that ends there
                        defk[b, per * i:per * (i + 1)] = o["defk"][per * i:per * (i + 1)]
                    obj[b, 0] += o["xend"][2, i]                         # quadrature objective: sum of the intervals
                    grad[b, blocks[i]:blocks[i + 1]] = Jx[O.nx * i + 2, blocks[i]:blocks[i + 1]]
            return dict(xend=xend, defects_kind=defk, objective=obj, grad=grad,
                        status=np.zeros(len(ps_rows), dtype=np.int32))

        rng = np.random.default_rng(99)
        ps = O.default_ps()[None, :] + 0.05 * rng.standard_normal((B, O.nvars))
        res = sharding.IntervalShardedEvaluator(evaluate_range, O.I)(ps)
        full = evaluate_range(ps, 1, O.I)
        ref = O.eval(ps[0], K=1)
        q.put((rank, res["_shard"].start, res["_shard"].stop,
               max(float(np.max(np.abs(res[k] - full[k]))) for k in ("xend", "defects_kind", "objective", "grad")),
               float(np.max(np.abs(res["defects_kind"][0] - ref["defk"]))), float(abs(res["objective"][0, 0] - ref["last"][2])),
               res["status"].dtype == np.int32))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_interval_sharding_assembles_by_summation():
    """B = 1 < world size: intervals are split across the ranks, one all-reduce(SUM) assembles every output"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_interval_worker, args=(r, 2, port, 1, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(o[1], o[2]) for o in out] == [(0, 3), (3, 6)]
    for o in out:
        assert o[3] <= 1e-15 and o[4] <= 1e-15 and o[5] <= 1e-13 and o[6]
