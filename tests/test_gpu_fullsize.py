"""BASELINE.json configs at their FULL sizes on the GPU, checked through size-independent properties and
against the oracle on a random sample of candidates (the oracle cannot do the whole batch in seconds).

    config 2  lqr_ms100_b4096      100 intervals x 4096 candidates, 8 sensitivity directions
    config 3  lotka_oed_64x16k     64 intervals x 16384 parameter samples, Fisher information accumulation
    config 4  dense20_256x65k      256 intervals, full 20 x 24 sensitivity Jacobian (a 1280-candidate chunk of the 65k)
"""
import numpy as np
import pytest

import problems as P

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b))) / max(1.0, float(np.max(np.abs(b)))) if a.size else 0.0


def test_config2_lqr_full_size():
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    from oracle import pyoracle as po
    import bench
    rng = np.random.default_rng(20261019 + 2)
    spec = P.lqr_spec(100, rng)
    lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(2))
    O = po.OracleLayer(spec)
    B = 4096
    ps = bench.lqr_inputs(N.nvars, cb.to_vec(ps_t), B, rng)
    want = nat.WANT_XEND | nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD
    r = N.eval_host(ps, want, objective_row=2, with_sums=True)
    r2 = N.eval_host(ps, want, objective_row=2, with_sums=True)
    for k in ("xend", "sens", "defects_kind", "defects_stage", "objective", "grad"):
        assert np.array_equal(r[k], r2[k]), f"{k} is not reproducible run to run"
        assert np.all(np.isfinite(r[k]))
    # sums over candidates (atomics: order varies, value within rounding)
    assert rel(r["objective_sum"], r["objective"].sum()) <= 1e-12
    assert rel(r["grad_sum"], r["grad"].sum(axis=0)) <= 1e-12
    # both orderings hold the same multiset of defects
    assert np.array_equal(np.sort(r["defects_kind"], axis=1), np.sort(r["defects_stage"], axis=1))
    # oracle on a random sample of candidates
    rows, cols, src = N.jac_structure()
    for b in rng.choice(B, 6, replace=False):
        o = O.eval(ps[b], K=2)
        assert rel(r["xend"][b].reshape(O.I, O.nx).T, o["xend"]) <= RTOL
        assert rel(r["defects_kind"][b], o["defk"]) <= RTOL
        assert rel(r["defects_stage"][b], o["defs"]) <= RTOL
        assert rel(r["objective"][b, 0], o["last"][1]) <= RTOL
    b = int(rng.integers(B))
    Jx, Jd, Jl = O.jacobians(ps[b], K=2)
    assert rel(r["grad"][b], Jl[1]) <= RTOL
    J = np.zeros((O.ndef, O.nvars))
    vals = np.where(src >= 0, r["sens"][b][np.maximum(src, 0)], np.where(src == -1, 1.0, -1.0))
    np.add.at(J, (rows - 1, cols - 1), vals)
    assert rel(J, Jd) <= RTOL
    # closed form of the linear dynamics over one control piece chain: d x_end / d x_0 = exp(a * 0.12) up to the
    # Tsit5 truncation error (h = 0.015: ~1e-13)
    a = ps[:, 1 + 0]  # interval 1: (p = (a, b), controls...) -> first variable is a
    sens = r["sens"].reshape(B, -1)
    blocks = N.block_structure()
    i = 7
    ns_i = blocks[i + 1] - blocks[i]
    off = sum(2 * (blocks[k + 1] - blocks[k]) for k in range(i))
    S = sens[:, off:off + 2 * ns_i].reshape(B, ns_i, 2)
    a_i = ps[:, blocks[i] + 2]
    assert np.allclose(S[:, 0, 0], np.exp(a_i * 0.12), rtol=1e-11, atol=0)


def test_config3_lotka_oed_full_size():
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    from oracle import pyoracle as po
    rng = np.random.default_rng(20261019 + 3)
    spec = P.config3_lotka_oed(64, rng)
    lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(2))
    O = po.OracleLayer(spec)
    B = 16384
    ps0 = cb.to_vec(ps_t)
    ps = np.tile(ps0, (B, 1))
    blocks = N.block_structure()
    # parameter samples (p2, p3) ~ U[0.8, 1.2]^2, the same in every interval of a candidate (parameter matchings = 0);
    # node states perturbed so that the state matchings are non-trivial
    pp = rng.uniform(0.8, 1.2, (B, 2))
    for i in range(64):
        nu0 = 0 if i == 0 else 9
        ps[:, blocks[i] + nu0: blocks[i] + nu0 + 2] = pp
        if i:
            ps[:, blocks[i]: blocks[i] + 2] += 0.05 * rng.standard_normal((B, 2))
    Finit = np.array([[0.3, 0.1], [0.1, 0.2]])
    r = N.eval_host(ps, nat.WANT_XEND | nat.WANT_FIM | nat.WANT_DEFECTS, fim_init=Finit)
    F = r["fim"].reshape(B, 2, 2)
    assert np.all(np.isfinite(F))
    assert np.array_equal(F[:, 0, 1], F[:, 1, 0])                     # symmetric by construction
    # F - F_init - F_node(last interval) is the integral of a Gram matrix over the last interval: PSD
    Fnode = ps[:, blocks[63] + 6: blocks[63] + 9]
    D = F - Finit[None] - np.stack([Fnode[:, 0], Fnode[:, 1], Fnode[:, 1], Fnode[:, 2]], axis=1).reshape(B, 2, 2)
    assert np.all(np.linalg.eigvalsh(D) >= -1e-13)
    assert rel(r["fim_sum"], r["fim"].sum(axis=0)) <= 1e-12          # sum over the 16k samples
    # parameter matchings vanish identically (same p in every interval)
    nd = N.n_defects
    for b in rng.choice(B, 5, replace=False):
        o = O.eval(ps[b], K=2)
        Fo = o["xend"][6:9, -1]
        assert rel(r["fim"][b], np.array([Fo[0], Fo[1], Fo[1], Fo[2]]) + Finit.ravel()) <= RTOL
        assert rel(r["defects_kind"][b], o["defk"]) <= RTOL
        assert rel(r["xend"][b].reshape(O.I, O.nx).T, o["xend"]) <= RTOL
    assert nd == 63 * (9 + 2)


def test_config4_dense20_full_intervals():
    """256 intervals x 1280 candidates (327 680 units, 0.16 G sensitivity values) through the tensor-path kernel:
    oracle on a sample, semigroup property of the sensitivities, reproducibility."""
    torch = pytest.importorskip("torch")
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    from oracle import pyoracle as po
    rng = np.random.default_rng(20261019 + 4)
    spec = P.config4_dense20(256, rng)
    lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(2))
    O = po.OracleLayer(spec)
    B, nvars, I, nx = 1280, N.nvars, 256, 20
    ps0 = cb.to_vec(ps_t)
    ps = np.tile(ps0, (B, 1)) + 0.1 * rng.standard_normal((B, nvars))
    ps = N.forward_init(ps)                      # continuous trajectories: nodes from the device forward sweep
    dev = torch.device("cuda:0")
    d_ps = torch.from_numpy(np.ascontiguousarray(ps.T)).to(dev)
    sz = N.out_sizes()
    outs = {f: torch.empty((sz[f], B), dtype=torch.float64, device=dev) for f in ("xend", "sens", "defects_kind")}
    want = nat.WANT_XEND | nat.WANT_SENS | nat.WANT_DEFECTS
    flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA
    N.set_stream(torch.cuda.current_stream().cuda_stream)
    N.eval_raw(d_ps.data_ptr(), B, B, want, flags, {k: v.data_ptr() for k, v in outs.items()}, 1)
    torch.cuda.synchronize()
    first = {k: v.clone() for k, v in outs.items()}
    N.eval_raw(d_ps.data_ptr(), B, B, want, flags, {k: v.data_ptr() for k, v in outs.items()}, 1)
    torch.cuda.synchronize()
    for k in outs:
        assert torch.equal(first[k], outs[k]), f"{k} is not reproducible"
    xend = outs["xend"].cpu().numpy().T.reshape(B, I, nx)
    defk = outs["defects_kind"].cpu().numpy().T
    assert np.all(np.isfinite(xend))
    # forward-initialised candidates: every state matching vanishes (test/core/multiple_shooting.jl:131-152)
    assert np.max(np.abs(defk)) <= 1e-11
    blocks = N.block_structure()
    for b in rng.choice(B, 3, replace=False):
        o = O.eval(ps[b], K=2)
        assert rel(xend[b].T, o["xend"]) <= RTOL
        assert rel(defk[b], o["defk"]) <= 1e-9      # defects ~ 0: absolute scale
    # sensitivities of one candidate against the oracle's dense Jacobian (intervals 0, 1, 128, 255)
    b = int(rng.integers(B))
    sens_b = outs["sens"][:, b].cpu().numpy()
    Jx, _, _ = O.jacobians(ps[b], K=2)
    off = np.concatenate([[0], np.cumsum([nx * (blocks[i + 1] - blocks[i]) for i in range(I)])])
    for i in (0, 1, 128, 255):
        ns = blocks[i + 1] - blocks[i]
        S = sens_b[off[i]:off[i + 1]].reshape(ns, nx).T
        assert rel(S, Jx[nx * i:nx * (i + 1), blocks[i]:blocks[i + 1]]) <= RTOL
