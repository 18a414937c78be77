"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical fixed-step inputs.

Tolerance (BASELINE.json north_star): <= 1e-10 relative in FP64 for states, defects, objective and
sensitivities; index-derived quantities (orderings, structure) bit-exact.
"""
import numpy as np
import pytest

import problems as P

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def rel_err(a, b):
    """max |a - b| / max(1, max |b|); non-finite entries must coincide (inf where they do not: never swallowed)"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    if not a.size:
        return 0.0
    if not (np.isfinite(a).all() and np.isfinite(b).all()):
        return float("inf")
    scale = max(1.0, float(np.max(np.abs(b))))
    return float(np.max(np.abs(a - b))) / scale


def batch_ps(layer_o, B, rng, scale=0.05):
    ps0 = layer_o.default_ps()
    return ps0[None, :] + scale * rng.standard_normal((B, layer_o.nvars))


def full_compare(spec, alg, B, rng, objective_row, check_traj=True, ps=None):
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    from oracle import pyoracle as po

    O = po.OracleLayer(spec)
    lay = P.product_layer(spec, alg)
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    method = {0: "tsit5", 1: "rk4", 2: "tsit5_adaptive"}[alg.method_id]
    if alg.method_id == 2:   # the problem's tolerances (ODEProblem kwargs), OrdinaryDiffEq's defaults otherwise
        po.set_tolerances(spec.get("kwargs", {}).get("abstol", 1e-6), spec.get("kwargs", {}).get("reltol", 1e-3))
    K = alg.steps_per_piece
    assert N.nvars == O.nvars and N.n_defects == O.ndef and N.n_samples == O.nsamples and N.n_row == O.nrow
    assert np.array_equal(N.block_structure(), O.block_structure())
    ps = batch_ps(O, B, rng) if ps is None else ps
    want = (nat.WANT_XEND | nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_JAC
            | nat.WANT_GRAD_COMPACT)
    if check_traj:
        want |= nat.WANT_TRAJ
    r = N.eval_host(ps, want, objective_row=objective_row)
    blocks = O.block_structure()
    nx, I = O.nx, O.I
    errs = dict(xend=0.0, defk=0.0, defs=0.0, obj=0.0, traj=0.0, sens=0.0, grad=0.0)
    rows, cols, src = N.jac_structure()
    g0, gn = N.grad_structure(objective_row)
    assert g0 + gn == N.nvars and np.array_equal(r["grad_compact"], r["grad"][:, g0:])
    assert not np.any(r["grad"][:, :g0])          # what the compact form leaves out is structurally zero
    for b in range(B):
        o = O.eval(ps[b], method=method, K=K)
        errs["xend"] = max(errs["xend"], rel_err(r["xend"][b].reshape(I, nx).T, o["xend"]))
        errs["defk"] = max(errs["defk"], rel_err(r["defects_kind"][b], o["defk"]))
        errs["defs"] = max(errs["defs"], rel_err(r["defects_stage"][b], o["defs"]))
        errs["obj"] = max(errs["obj"], rel_err(r["objective"][b], o["last"][objective_row - 1]))
        if check_traj:
            errs["traj"] = max(errs["traj"], rel_err(r["traj"][b].reshape(O.nsamples, O.nrow).T, o["u"]))
        if b < 3:  # dense oracle Jacobians are the slow part
            Jx, Jd, Jl = O.jacobians(ps[b], method=method, K=K)
            off = 0
            for i in range(I):
                ns = blocks[i + 1] - blocks[i]
                S = r["sens"][b][off:off + nx * ns].reshape(ns, nx).T          # [k, d]
                errs["sens"] = max(errs["sens"], rel_err(S, Jx[nx * i:nx * (i + 1), blocks[i]:blocks[i + 1]]))
                off += nx * ns
            errs["grad"] = max(errs["grad"], rel_err(r["grad"][b], Jl[objective_row - 1]))
            # Jacobian of the kind-major constraints assembled from the structure + sens
            J = np.zeros((O.ndef, O.nvars))
            vals = np.where(src >= 0, r["sens"][b][np.maximum(src, 0)], np.where(src == -1, 1.0, -1.0))
            np.add.at(J, (rows - 1, cols - 1), vals)
            errs["jac"] = max(errs.get("jac", 0.0), rel_err(J, Jd))
            # the same values assembled on the device (K5, WANT_JAC): bit-identical to the gather from sens
            assert np.all(src[:N.jac_nnz_var] >= 0) and np.all(src[N.jac_nnz_var:] < 0)
            assert np.array_equal(r["jac_vals"][b], vals[:N.jac_nnz_var]), "device cons_j values differ from sens gather"
    return errs


def assert_ok(errs):
    bad = {k: v for k, v in errs.items() if not v <= RTOL}
    assert not bad, f"parity errors above {RTOL}: {bad} (all: {errs})"


def test_lotka_ms4_reference_golden():
    """G1: the reference's own 15-element defect vector (test/core/multiple_shooting.jl:54-72, atol 1e-10)."""
    import corleone_b200 as cb
    lay = P.product_layer(P.lotka_spec(), cb.Tsit5(8))
    ps, st = cb.setup(None, lay)
    traj, _ = lay(None, ps, st)
    G1 = np.array([1.375754960969482, 0.22357357513551113, 1.2417260108009396] * 3 + [0.0] * 6)
    assert cb.is_shooting_solution(traj)
    assert np.allclose(cb.shooting_constraints(traj), G1, atol=1e-10, rtol=0)
    assert len(cb.shooting_constraints(traj)) == cb.get_number_of_shooting_constraints(lay, ps, st) == 15
    assert traj.u.shape == (121, 4) and len(np.unique(traj.t)) == 121
    assert list(traj.shooting_indices) == [31, 62, 93]
    # MS objective golden (test/examples/lotka_ms.jl:45)
    assert abs(traj["x₃"][-1] - 1.2417260108009376) < 1e-9


@pytest.mark.parametrize("method", ["tsit5", "rk4"])
def test_lotka_ms4_parity(method):
    import corleone_b200 as cb
    alg = cb.Tsit5(2) if method == "tsit5" else cb.RK4(3)
    rng = np.random.default_rng(20261019)
    assert_ok(full_compare(P.lotka_spec(controls=rng.uniform(0, 1, 120)), alg, 5, rng, objective_row=3))


def test_lotka_ms4_quadrature_parity():
    import corleone_b200 as cb
    rng = np.random.default_rng(7)
    spec = P.lotka_spec(controls=rng.uniform(0, 1, 120), quadrature=True)
    assert_ok(full_compare(spec, cb.Tsit5(2), 4, rng, objective_row=3))


def test_lotka_ss_chunked_parity():
    """single shooting over 120 pieces (> subdivide = 100, so the reference chunks the tables)"""
    import corleone_b200 as cb
    rng = np.random.default_rng(3)
    spec = P.lotka_spec(shooting_points=None, controls=rng.uniform(0, 1, 120), tunable_ic=[2, 1])
    assert_ok(full_compare(spec, cb.Tsit5(1), 3, rng, objective_row=3))


def test_lotka_ms24_config1():
    import corleone_b200 as cb
    rng = np.random.default_rng(20261019 + 1)
    assert_ok(full_compare(P.lotka_ms24_spec(rng), cb.Tsit5(4), 1, rng, objective_row=3))


def test_lqr_config2_small_batch():
    import corleone_b200 as cb
    rng = np.random.default_rng(20261019 + 2)
    assert_ok(full_compare(P.lqr_spec(100, rng), cb.Tsit5(2), 33, rng, objective_row=2, check_traj=True))


def test_egerstedt_three_controls_permuted():
    import corleone_b200 as cb
    rng = np.random.default_rng(5)
    assert_ok(full_compare(P.egerstedt_spec(), cb.Tsit5(2), 3, rng, objective_row=3))


def test_dense20_parity_small():
    import corleone_b200 as cb
    rng = np.random.default_rng(20261019 + 4)
    spec = P.dense20_spec(4, rng)
    assert_ok(full_compare(spec, cb.Tsit5(2), 3, rng, objective_row=1, check_traj=False))


@pytest.mark.parametrize("nx", [16, 24, 32])
@pytest.mark.parametrize("method", ["tsit5", "rk4"])
def test_dense_family_on_the_tensor_path(nx, method):
    """the dense family x' = W x - x^3 + b u at the other sizes the FP64 tensor-path kernel is instantiated for (lane
    ownership S = nx/4 states: 4 -- no padding, 6, 8; 12 and 8 warps per block for 24 and 32): full parity, a batch that
    is not a multiple of the 8-candidate item, more items than SMs"""
    import corleone_b200 as cb
    rng = np.random.default_rng(20261019 + nx)
    spec = P.dense20_spec(5, rng, nx=nx)
    alg = cb.Tsit5(2) if method == "tsit5" else cb.RK4(2)
    assert_ok(full_compare(spec, alg, 3, rng, objective_row=2, check_traj=True))
    assert_ok(full_compare(spec, alg, 261, rng, objective_row=nx, check_traj=False))


@pytest.mark.parametrize("nx,alg_name,K", [(32, "rk4", 5), (32, "tsit5", 5), (24, "tsit5", 4), (20, "tsit5", 8)])
def test_dense_family_with_many_stages_per_interval(nx, alg_name, K):
    """layers whose stage count does not fit three ring slots of the tile-queue kernel: two slots, one slot (no look-ahead),
    and beyond that the lockstep kernel -- the same results"""
    import corleone_b200 as cb
    rng = np.random.default_rng(7 * nx + K)
    spec = P.dense20_spec(3, rng, nx=nx)
    alg = cb.Tsit5(K) if alg_name == "tsit5" else cb.RK4(K)
    assert_ok(full_compare(spec, alg, 19, rng, objective_row=1, check_traj=True))


@pytest.mark.parametrize("model", ["lotka3", "lqr"])
@pytest.mark.parametrize("B", [1, 2, 7, 31])
def test_small_evaluation_in_one_kernel_equals_the_pipeline(model, B):
    """a few candidates through host pointers: the shooting kernel's last block computes the objective and copies the result
    block out (small_tail) -- bit for bit what shooting kernel + objective kernel + copy node deliver, on repeated calls
    (the captured graph is replayed) and with a quadrature row as objective (ordered sum over the intervals)"""
    import os
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(B)
    spec = P.lotka_spec(controls=rng.uniform(0, 1, 120), quadrature=True) if model == "lotka3" else P.lqr_spec(7, rng)
    spec = dict(spec, quadrature_indices=[3] if model == "lotka3" else [2])
    row = 3 if model == "lotka3" else 2
    lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(2))
    ps = cb.to_vec(ps_t)[None, :] + 0.05 * rng.standard_normal((B, N.nvars))
    psT = np.ascontiguousarray(ps.T)                       # variable-major: the small path
    want = nat.WANT_JAC | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_XEND
    sz = N.out_sizes()
    fields = ("jac_vals", "defects_kind", "objective", "grad", "xend")

    def run():
        out = {f: np.full((sz[f], B), np.nan) for f in fields}
        for _ in range(3):
            N.eval_raw(psT.ctypes.data, B, B, want, nat.PS_SOA | nat.OUT_SOA, {f: out[f].ctypes.data for f in fields}, row)
        return out

    fused = run()
    os.environ["CB200_NO_SMALL_TAIL"] = "1"
    try:
        lay2, ps_t2, st2, N2 = P.native_from_spec(spec, cb.Tsit5(2))
        N_keep, N = N, N2
        plain = run()
    finally:
        del os.environ["CB200_NO_SMALL_TAIL"]
    for f in fields:
        assert np.array_equal(fused[f], plain[f]), f
        assert np.isfinite(fused[f]).all()
    assert N_keep.launch_count() < N2.launch_count()       # one kernel per call instead of two


def test_lotka_oed_fim_parity():
    """states-only integration of [x; G; F_triu]; FIM read-out (oed.jl:388-398) and its sum over candidates"""
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    from oracle import pyoracle as po
    rng = np.random.default_rng(11)
    for sp in (None, [0.0, 3.0, 6.0, 9.0]):
        spec = P.lotka_oed_spec(shooting_points=sp, fishing=rng.uniform(0, 1, 48))
        O = po.OracleLayer(spec)
        lay = P.product_layer(spec, cb.Tsit5(2))
        ps_t, st = cb.setup(None, lay)
        (st if sp is None else st["interval_1"])["_fim"] = (7, 2)
        N = lay._native(st, ps_t)
        B = 37
        ps = batch_ps(O, B, rng, 0.02)
        r = N.eval_host(ps, nat.WANT_XEND | nat.WANT_FIM | nat.WANT_DEFECTS)
        Fsum = np.zeros(4)
        for b in range(B):
            o = O.eval(ps[b], K=2)
            F = o["xend"][6:9, -1]
            Fm = np.array([F[0], F[1], F[1], F[2]])
            assert rel_err(r["fim"][b], Fm) <= RTOL
            assert rel_err(r["defects_kind"][b], o["defk"]) <= RTOL
            Fsum += Fm
        assert rel_err(r["fim_sum"], Fsum) <= 1e-12 * B + RTOL


def test_empty_batch_and_errors():
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    lay = P.product_layer(P.lotka_spec(), cb.Tsit5(1))
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    r = N.eval_host(np.zeros((0, N.nvars)), nat.WANT_XEND)
    assert r["xend"].shape == (0, 12)
    with pytest.raises(nat.NativeError):
        N.eval_host(np.zeros((1, N.nvars)), nat.WANT_OBJ, objective_row=99)
    with pytest.raises(nat.NativeError):
        N.eval_host(np.zeros((1, N.nvars)), nat.WANT_FIM)
    # failure flags: a candidate whose node is NaN is reported, its neighbours are not
    ps = np.tile(cb.to_vec(ps_t), (5, 1))
    ps[3, 40] = np.nan
    ps[1, 2] = 1e308          # overflows during integration
    r = N.eval_host(ps, nat.WANT_XEND | nat.WANT_DEFECTS)
    assert list(r["status"]) == [0, 1, 0, 1, 0]
    assert np.all(np.isfinite(r["xend"][[0, 2, 4]]))


def test_device_resident_soa_path():
    """device pointers, SoA in and out (the layout the kernels use natively) == host candidate-major path"""
    torch = pytest.importorskip("torch")
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(2)
    spec = P.lqr_spec(100, rng)
    lay = P.product_layer(spec, cb.Tsit5(2))
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    B = 257
    ps = cb.to_vec(ps_t)[None, :] + 0.1 * rng.standard_normal((B, N.nvars))
    want = nat.WANT_XEND | nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD
    ref = N.eval_host(ps, want, objective_row=2)
    dev = torch.device("cuda:0")
    d_ps = torch.from_numpy(np.ascontiguousarray(ps.T)).to(dev)           # [nvars, B]
    sz = N.out_sizes()
    outs = {k: torch.empty((sz[k], B), dtype=torch.float64, device=dev)
            for k in ("xend", "sens", "defects_kind", "defects_stage", "objective", "grad")}
    N.set_stream(torch.cuda.current_stream().cuda_stream)
    N.eval_raw(d_ps.data_ptr(), B, B, want, nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA,
               {k: v.data_ptr() for k, v in outs.items()}, objective_row=2)
    torch.cuda.synchronize()
    for k, v in outs.items():
        assert np.array_equal(v.cpu().numpy().T, ref[k]), k
    assert N.last_kernel_ms() > 0


def test_forward_initialization_batch_vs_oracle():
    """cb200_forward_init (device sweep over intervals) == oracle restatement of node_initialization.jl:152-170,
    incl. fixed_indices; defects of a forward-initialised candidate vanish (test/core/multiple_shooting.jl:131-152)."""
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    from oracle import pyoracle as po
    rng = np.random.default_rng(17)
    for spec, K in ((P.lotka_spec(controls=rng.uniform(0, 1, 120), quadrature=True), 2),
                    (P.lotka_spec(controls=rng.uniform(0, 1, 120)), 2),
                    (P.lqr_spec(20, rng), 2)):
        O = po.OracleLayer(spec)
        lay = P.product_layer(spec, cb.Tsit5(K))
        ps_t, st = cb.setup(None, lay)
        N = lay._native(st, ps_t)
        B = 7
        ps = batch_ps(O, B, rng)
        for fixed in ((), (3,)):
            got = N.forward_init(ps, fixed)
            for b in range(B):
                ref = O.forward_init(ps[b], fixed, K=K)
                assert rel_err(got[b], ref) <= RTOL
            if not fixed:
                r = N.eval_host(got, nat.WANT_DEFECTS)
                nstate = sum(int(np.sum(st[k]["shooting_indices"] <= O.nx)) for k in list(st)[1:])
                if not spec["quadrature_indices"]:
                    assert np.max(np.abs(r["defects_kind"][:, :nstate])) <= 1e-12
    # the reference-API entry point (one candidate) goes through the same kernel
    lay = P.product_layer(P.lotka_spec(), cb.Tsit5(8), initialization=cb.forward_initialization)
    ps_f = lay.initialparameters(None)
    assert np.all(np.isfinite(cb.to_vec(ps_f)))


def test_oed_criteria_batched():
    """the six criteria of criteria.jl:33-63 per candidate on the device vs NumPy on the same FIM"""
    import corleone_b200 as cb
    rng = np.random.default_rng(23)
    spec = P.lotka_oed_spec(shooting_points=[0.0, 3.0, 6.0, 9.0], fishing=rng.uniform(0, 1, 48))
    lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(2))
    B = 65
    ps = cb.to_vec(ps_t)[None, :] + 0.02 * rng.standard_normal((B, N.nvars))
    Finit = np.array([[0.5, 0.1], [0.1, 0.4]])
    crit, F = cb.batched_criteria(N, ps, fim_init=Finit)
    host = [c() for c in (cb.ACriterion, cb.DCriterion, cb.ECriterion, cb.FisherACriterion, cb.FisherDCriterion,
                          cb.FisherECriterion)]
    for b in range(B):
        ref = np.array([c(F[b]) for c in host])
        assert np.allclose(crit[b], ref, rtol=1e-10, atol=0), (b, crit[b], ref)


@pytest.mark.parametrize("method", ["tsit5", "rk4"])
def test_dense20_tensor_path_many_groups(method):
    """dense 20-state model on the FP64 tensor-path kernel with more unit groups than SMs (several items per
    persistent block, both shared-memory item buffers reused) and a batch that is not a multiple of the group size"""
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    from oracle import pyoracle as po
    rng = np.random.default_rng(20261019 + 44)
    spec = P.dense20_spec(8, rng)
    alg = cb.Tsit5(2) if method == "tsit5" else cb.RK4(2)
    O = po.OracleLayer(spec)
    lay = P.product_layer(spec, alg)
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    B = 203
    ps = batch_ps(O, B, rng, 0.1)
    want = nat.WANT_XEND | nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_JAC | nat.WANT_TRAJ
    r = N.eval_host(ps, want, objective_row=1, with_sums=True)
    assert rel_err(r["grad_sum"], r["grad"].sum(axis=0)) <= 1e-12
    rows, cols, src = N.jac_structure()
    blocks = O.block_structure()
    nx, I = O.nx, O.I
    for b in (0, 4, 5, 101, 199, 202):
        o = O.eval(ps[b], method=method, K=2)
        assert rel_err(r["xend"][b].reshape(I, nx).T, o["xend"]) <= RTOL
        assert rel_err(r["defects_kind"][b], o["defk"]) <= RTOL
        assert rel_err(r["defects_stage"][b], o["defs"]) <= RTOL
        assert rel_err(r["traj"][b].reshape(O.nsamples, O.nrow).T, o["u"]) <= RTOL
        assert rel_err(r["objective"][b], o["last"][0]) <= RTOL
    for b in (5, 202):
        Jx, Jd, Jl = O.jacobians(ps[b], method=method, K=2)
        off = 0
        for i in range(I):
            ns = blocks[i + 1] - blocks[i]
            S = r["sens"][b][off:off + nx * ns].reshape(ns, nx).T
            assert rel_err(S, Jx[nx * i:nx * (i + 1), blocks[i]:blocks[i + 1]]) <= RTOL
            off += nx * ns
        assert rel_err(r["grad"][b], Jl[0]) <= RTOL
        J = np.zeros((O.ndef, O.nvars))
        np.add.at(J, (rows - 1, cols - 1), np.concatenate([r["jac_vals"][b], np.where(src[N.jac_nnz_var:] == -1, 1.0, -1.0)]))
        assert rel_err(J, Jd) <= RTOL


def test_offgrid_nodes_control_matchings_and_tunable_ic():
    """shooting nodes that are not control grid points give control matchings (src/single_shooting.jl:333-340,
    src/multiple_shooting.jl:150-163) and non-uniform piece lengths (per-piece step sizes); tunable initial
    conditions on the first interval (test/core/multiple_shooting.jl style)"""
    import corleone_b200 as cb
    rng = np.random.default_rng(31)
    spec = P.lotka_spec(shooting_points=[0.0, 2.95, 6.0, 8.17], controls=rng.uniform(0, 1, 120), tunable_ic=[2, 1])
    assert_ok(full_compare(spec, cb.Tsit5(2), 4, rng, objective_row=3))
    spec = P.lotka_spec(shooting_points=[0.0, 2.95, 6.0, 8.17], controls=rng.uniform(0, 1, 120), quadrature=True)
    assert_ok(full_compare(spec, cb.RK4(2), 35, rng, objective_row=3))


def test_layer_without_controls():
    """no ControlParameter at all: index_grid = control_indices = [], one piece = the problem's tspan
    (src/single_shooting.jl:329-332); every parameter is tunable"""
    import corleone_b200 as cb
    rng = np.random.default_rng(37)
    spec = P.lotka_spec(shooting_points=None, tunable_ic=[1, 2])
    spec["controls"] = []
    assert_ok(full_compare(spec, cb.Tsit5(40), 3, rng, objective_row=3))


def test_oed_second_order_sensitivities():
    """sensitivities of the augmented OED system [x; G; F_triu] w.r.t. nodes, parameters, the control and the
    measurement weights (what ForwardDiff through the OED layer yields, lib/CorleoneOED/src/augmentation.jl:131-253):
    d F(t_end) / d(everything) is what a design optimiser differentiates the criteria with."""
    import corleone_b200 as cb
    rng = np.random.default_rng(41)
    for sp in (None, [0.0, 3.0, 6.0, 9.0]):
        spec = P.lotka_oed_spec(shooting_points=sp, fishing=rng.uniform(0, 1, 48), w=0.7)
        assert_ok(full_compare(spec, cb.Tsit5(2), 3, rng, objective_row=7, check_traj=False))   # objective = F11
    spec = P.config3_lotka_oed(6, rng)
    assert_ok(full_compare(spec, cb.RK4(2), 34, rng, objective_row=9, check_traj=False))          # F22
    assert_ok(full_compare(P.lin1d_oed_spec(), cb.Tsit5(1), 3, rng, objective_row=3, check_traj=False))


def test_oed_criteria_gradient_vs_finite_differences():
    import corleone_b200 as cb
    rng = np.random.default_rng(43)
    spec = P.lotka_oed_spec(shooting_points=None, fishing=rng.uniform(0, 1, 48), w=0.8)
    lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(2))
    ps = cb.to_vec(ps_t)[None, :] + 0.02 * rng.standard_normal((2, N.nvars))
    Finit = 0.3 * np.eye(2)
    g = cb.criteria_gradient(N, ps, fim_init=Finit)
    h = 1e-6
    for v in rng.choice(N.nvars, 12, replace=False):
        pp, pm = ps.copy(), ps.copy()
        pp[:, v] += h
        pm[:, v] -= h
        cp, _ = cb.batched_criteria(N, pp, fim_init=Finit)
        cm, _ = cb.batched_criteria(N, pm, fim_init=Finit)
        fd = (cp - cm) / (2 * h)
        assert np.allclose(g[:, :, v], fd, rtol=2e-6, atol=1e-7), (v, g[:, :, v], fd)


@pytest.mark.parametrize("name", ["vdp", "cartpole", "rocket", "quadrotor", "threetank", "batchreactor", "lotkacomp", "f8",
                                  "doubleosc"])
def test_benchmark_suite_models_forward_mode_ad(name):
    """right-hand sides of lib/OptimalControlBenchmarks written once on a generic scalar; the kernels take the
    directional derivatives by forward-mode AD on the device (sin, cos, exp, division covered)"""
    import corleone_b200 as cb
    rng = np.random.default_rng(47)
    spec, row, scale = {"vdp": (P.vdp_spec(5, rng), 3, 0.05), "cartpole": (P.cartpole_spec(4, rng), 5, 0.05),
                        "rocket": (P.rocket_spec(4, rng), 3, 0.001), "quadrotor": (P.quadrotor_spec(3, rng), 7, 0.05),
                        "threetank": (P.threetank_spec(4, rng), 4, 0.02),
                        "batchreactor": (P.batchreactor_spec(4, rng), 2, 0.01),
                        "lotkacomp": (P.lotkacomp_spec(4, rng), 3, 0.05), "f8": (P.f8_spec(4, rng), 4, 0.02),
                        "doubleosc": (P.doubleosc_spec(3, rng), 6, 0.05)}[name]
    from oracle import pyoracle as po
    O = po.OracleLayer(spec)
    ps = O.default_ps()[None, :] + scale * rng.standard_normal((33, O.nvars))
    assert_ok(full_compare(spec, cb.Tsit5(3), 33, rng, objective_row=row, ps=ps))
    assert_ok(full_compare(spec, cb.RK4(4), 3, rng, objective_row=row, ps=ps[:3]))


@pytest.mark.parametrize("name", sorted(P.BENCH_TABLE))
def test_benchmark_suite_remaining_problems(name):
    """the rest of lib/OptimalControlBenchmarks/src/problems (every problem file has a compiled-in right-hand side):
    states, defects, objective, sensitivities, gradient and constraint Jacobian against the oracle"""
    import corleone_b200 as cb
    rng = np.random.default_rng(61)
    spec, row, scale = P.bench_spec(name, 3 if name == "ductedfan" else 4)
    from oracle import pyoracle as po
    O = po.OracleLayer(spec)
    ps = O.default_ps()[None, :] * (1.0 + scale * rng.standard_normal((9, O.nvars))) + scale * rng.standard_normal((9, O.nvars))
    assert_ok(full_compare(spec, cb.Tsit5(3), 9, rng, objective_row=row, ps=ps))
    assert_ok(full_compare(spec, cb.RK4(4), 2, rng, objective_row=row, ps=ps[:2]))


def test_every_memory_layout_of_cb200_eval_agrees():
    """host candidate-major (chunked, multi-stream; last chunk partial), host variable-major, device candidate-major
    and device variable-major calls give bit-identical results"""
    torch = pytest.importorskip("torch")
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(53)
    spec = P.lqr_spec(12, rng)
    lay = P.product_layer(spec, cb.Tsit5(2))
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    B, nvars = 1500, N.nvars
    ps = cb.to_vec(ps_t)[None, :] + 0.1 * rng.standard_normal((B, nvars))
    want = nat.WANT_XEND | nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_JAC
    fields = ("xend", "sens", "defects_kind", "defects_stage", "objective", "grad", "jac_vals")
    ref = N.eval_host(ps, want, objective_row=2)                                   # host, candidate-major, 6 chunks
    sz = N.out_sizes()
    # host, variable-major in and out
    ps_soa = np.ascontiguousarray(ps.T)
    out = {f: np.empty((sz[f], B)) for f in fields}
    N.eval_raw(ps_soa.ctypes.data, B, B, want, nat.PS_SOA | nat.OUT_SOA, {f: out[f].ctypes.data for f in fields}, 2)
    for f in fields:
        assert np.array_equal(out[f].T, ref[f]), ("host soa", f)
    dev = torch.device("cuda:0")
    N.set_stream(torch.cuda.current_stream().cuda_stream)
    # device, candidate-major in and out (Julia CuArray-style nvars x B matrices)
    d_ps = torch.from_numpy(ps).to(dev)
    d_out = {f: torch.empty((B, sz[f]), dtype=torch.float64, device=dev) for f in fields}
    N.eval_raw(d_ps.data_ptr(), B, nvars, want, nat.MEM_DEVICE, {f: d_out[f].data_ptr() for f in fields}, 2)
    torch.cuda.synchronize()
    for f in fields:
        assert np.array_equal(d_out[f].cpu().numpy(), ref[f]), ("device candidate-major", f)
    # device, variable-major with a padded leading dimension
    ldb = B + 12
    d_pad = torch.zeros((nvars, ldb), dtype=torch.float64, device=dev)
    d_pad[:, :B] = torch.from_numpy(ps_soa).to(dev)
    d_o2 = {f: torch.empty((sz[f], B), dtype=torch.float64, device=dev) for f in fields}
    N.eval_raw(d_pad.data_ptr(), B, ldb, want, nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA,
               {f: d_o2[f].data_ptr() for f in fields}, 2)
    torch.cuda.synchronize()
    for f in fields:
        assert np.array_equal(d_o2[f].cpu().numpy().T, ref[f]), ("device soa", f)


def test_reference_goldens_through_the_device_path():
    """the reference's own golden values (adaptive Tsit5) reproduced by the CUDA path at converged fixed steps:
    G4 hybrid/forward node (test/core/multiple_shooting.jl:202-222, atol 1e-10), G3 single-shooting objective
    (test/examples/lotka_oc.jl:80), G7 A-criterion single / multiple shooting (lib/CorleoneOED/test/core/lotka_oed.jl:124),
    G9 analytic 1-D OED (lib/CorleoneOED/test/core/1d_oed.jl:11-31)"""
    import json
    import os
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_goldens.json")))
    # G4: custom node on interval 2, forward sweep with fixed_indices = [2, 4]
    lay = P.product_layer(P.lotka_spec(), cb.Tsit5(8))
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    blocks = N.block_structure()
    ps = cb.to_vec(ps_t)
    ps[blocks[1]:blocks[1] + 3] = [0.5, 3.0, -5.0]
    out = N.forward_init(ps[None, :], fixed_indices=(2, 4))[0]
    assert np.allclose(out[blocks[2]:blocks[2] + 3], g["G4_hybrid_forward_node"]["values"], atol=1e-10, rtol=0)
    assert np.array_equal(out[blocks[3]:blocks[3] + 3], [0.5, 0.7, 0.0])
    # G3: single-shooting objective at the initial point
    lay = P.product_layer(P.lotka_spec(shooting_points=None), cb.Tsit5(4))
    ps_t, st = cb.setup(None, lay)
    r = lay._native(st, ps_t).eval_host(cb.to_vec(ps_t)[None, :], nat.WANT_OBJ, objective_row=3)
    assert abs(r["objective"][0, 0] - g["G3_lotka_ss_objective"]["value"]) < 1e-9
    # G7: A-criterion at the initial point
    for sp, key, tol in ((None, "single_shooting", 1.5e-8), ([0.0, 3.0, 6.0, 9.0], "multiple_shooting", 2e-7)):
        lay, ps_t, st, N = P.native_from_spec(P.lotka_oed_spec(shooting_points=sp), cb.Tsit5(8))
        crit, F = cb.batched_criteria(N, cb.to_vec(ps_t)[None, :])
        assert abs(crit[0, 0] / g["G7_oed_A_criterion"][key] - 1) < tol
    # G9: x' = p x, p = -2: F = int_0^1 t^2 e^{2 p t} dt
    lay, ps_t, st, N = P.native_from_spec(P.lin1d_oed_spec(), cb.Tsit5(4))
    r = N.eval_host(cb.to_vec(ps_t)[None, :], nat.WANT_FIM | nat.WANT_XEND)
    p = -2.0
    F_exact = (np.exp(2 * p) * (2 * p * p - 2 * p + 1) - 1) / (4 * p ** 3)
    assert abs(r["fim"][0, 0] - F_exact) < 1e-10 and abs(r["xend"][0, 0] - np.exp(p)) < 1e-12


@pytest.mark.parametrize("name", ["lotka_quad", "lqr", "dense20", "lotka_ss"])
def test_trajectory_sensitivities_for_point_constraints(name):
    """d(every saved trajectory sample)/d(ps): WANT_TRAJ_SENS + host assembly == the oracle's tangents through the
    merged trajectory (what the point constraints of src/dynprob.jl:30-58 differentiate)"""
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    from oracle import pyoracle as po
    rng = np.random.default_rng(59)
    spec, alg = {"lotka_quad": (P.lotka_spec(controls=rng.uniform(0, 1, 120), quadrature=True), cb.Tsit5(2)),
                 "lqr": (P.lqr_spec(7, rng), cb.RK4(2)),
                 "dense20": (P.dense20_spec(3, rng), cb.Tsit5(2)),
                 "lotka_ss": (P.lotka_spec(shooting_points=None, controls=rng.uniform(0, 1, 120), tunable_ic=[1, 2]),
                              cb.Tsit5(1))}[name]
    O = po.OracleLayer(spec)
    lay = P.product_layer(spec, alg)
    ps_t, st = cb.setup(None, lay)
    N = lay._native(st, ps_t)
    B = 4
    ps = batch_ps(O, B, rng)
    r = N.eval_host(ps, nat.WANT_TRAJ | nat.WANT_TRAJ_SENS | nat.WANT_SENS)
    method = "tsit5" if alg.method_id == 0 else "rk4"
    for b in (0, B - 1):
        u, Jref = O.traj_jacobian(ps[b], method=method, K=alg.steps_per_piece)
        assert rel_err(r["traj"][b].reshape(O.nsamples, O.nrow).T, u) <= RTOL
        J = cb.trajectory_jacobian(lay, st, ps_t, r, b)
        assert rel_err(J, Jref) <= RTOL


def test_dense20_layers_on_one_device_share_their_constant_block():
    """the thread-per-direction dense kernels read W, b from one __constant__ block per device: a second live layer with
    DIFFERENT data is refused (it would silently change the first one's values-only evaluations), identical data and
    sequential use are fine"""
    import gc
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    gc.collect()
    spec = P.dense20_spec(2)
    lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(2))
    lay2, ps2, st2, N2 = P.native_from_spec(P.dense20_spec(3), cb.Tsit5(2))        # same data: accepted
    other = dict(spec, model_data=np.asarray(spec["model_data"]) * 1.5)
    with pytest.raises(nat.NativeError, match="different model_data"):
        P.native_from_spec(other, cb.Tsit5(2))
    x = cb.to_vec(ps_t)[None, :]
    ref = N.eval_host(x, nat.WANT_XEND)["xend"].copy()
    N.close(); N2.close()
    for s in (st, st2):
        for v in (s.values() if "tspans" not in s else [s]):
            v.pop("_native", None)
    gc.collect()
    lay3, ps3, st3, N3 = P.native_from_spec(other, cb.Tsit5(2))                     # alone on the device: accepted
    assert not np.array_equal(N3.eval_host(x, nat.WANT_XEND)["xend"], ref)
    N3.close()
