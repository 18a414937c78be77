"""Adaptive Tsit5 on the device (cb200_method 2, the reference's default mode): parity with the oracle's restatement of
OrdinaryDiffEq's stepping (<= 1e-10; in practice the accepted step sequences coincide and the difference is rounding),
and the reference's golden vectors straight from the device."""
import numpy as np
import pytest

import problems as P
from test_gpu_parity import RTOL, assert_ok, full_compare, rel_err

pytestmark = pytest.mark.gpu

CORE = dict(abstol=1.0e-8, reltol=1.0e-6)


def _layer(spec, **kw):
    import corleone_b200 as cb
    lay = P.product_layer(spec, cb.Tsit5(adaptive=True), **kw)
    ps, st = cb.setup(None, lay)
    return lay, ps, st


def test_reference_goldens_from_the_device():
    """G1 (test/core/multiple_shooting.jl:54-72), G2/G3 (test/examples/lotka_ms.jl:45, lotka_oc.jl:80), G7
    (lib/CorleoneOED/test/core/lotka_oed.jl:124) -- to ~1e-14, where the fixed-step path reaches 5e-11 ... 7e-8"""
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    spec = dict(P.lotka_spec(), kwargs=CORE)
    lay, ps, st = _layer(spec)
    traj, _ = lay(None, ps, st)
    G1 = np.array([1.375754960969482, 0.22357357513551113, 1.2417260108009396,
                   1.3757549609694817, 0.22357357513551046, 1.2417260108009387,
                   1.3757549609694821, 0.2235735751355138, 1.2417260108009425] + [0.0] * 6)
    assert np.max(np.abs(cb.shooting_constraints(traj) - G1)) <= 2e-14
    assert abs(traj["x₃"][-1] - 1.2417260108009376) <= 2e-14
    lay, ps, st = _layer(dict(P.lotka_spec(shooting_points=None), kwargs=CORE))
    traj, _ = lay(None, ps, st)
    assert abs(traj["x₃"][-1] - 6.062277454291031) <= 5e-14
    for sp, gold in ((None, 0.05352869250783344), ([0.0, 3.0, 6.0, 9.0], 0.6199466255548527)):
        lay, ps_t, st, N = P.native_from_spec(P.lotka_oed_spec(shooting_points=sp), cb.Tsit5(adaptive=True))
        r = N.eval_host(cb.to_vec(ps_t)[None, :], nat.WANT_FIM | nat.WANT_CRIT)
        assert abs(r["criteria"][0, 0] / gold - 1.0) <= 1e-13, (sp, r["criteria"][0, 0])


def test_hybrid_initialisation_golden_from_the_device():
    """G4 (test/core/multiple_shooting.jl:202-222): custom node on interval 2, forward sweep with fixed_indices = [2, 4]
    -- the sweep runs in cb200_forward_init with adaptive steps"""
    import corleone_b200 as cb
    lay, ps_t, st = _layer(dict(P.lotka_spec(), kwargs=CORE))
    N = lay._native(st, ps_t)
    blocks = N.block_structure()
    ps = cb.to_vec(ps_t)
    ps[blocks[1]:blocks[1] + 3] = [0.5, 3.0, -5.0]
    out = N.forward_init(ps[None, :], fixed_indices=(2, 4))[0]
    assert np.max(np.abs(out[blocks[2]:blocks[2] + 3] - [0.2875960819428889, 0.27699894224421623, -1.088467272254529])) <= 2e-14
    assert np.array_equal(out[blocks[3]:blocks[3] + 3], [0.5, 0.7, 0.0])
    # forward initialisation leaves exactly zero defects (:131-152)
    from corleone_b200 import native as nat
    full = N.forward_init(cb.to_vec(ps_t)[None, :])
    assert not N.eval_host(full, nat.WANT_DEFECTS)["defects_kind"].any()


@pytest.mark.parametrize("case", ["lotka_ms", "lotka_ss", "lqr", "lotka2", "lin1d"])
def test_adaptive_parity_with_the_oracle(case):
    import corleone_b200 as cb
    rng = np.random.default_rng(71)
    alg = cb.Tsit5(adaptive=True)
    if case == "lotka_ms":
        spec, row = dict(P.lotka_spec(controls=rng.uniform(0, 1, 120), quadrature=True), kwargs=CORE), 3
    elif case == "lotka_ss":
        spec, row = dict(P.lotka_spec(shooting_points=None, controls=rng.uniform(0, 1, 120)), kwargs=CORE), 3
    elif case == "lqr":
        spec, row = dict(P.lqr_spec(12, rng), kwargs=dict(abstol=1e-9, reltol=1e-7)), 2
    elif case == "lotka2":
        spec, row = dict(model="lotka2", u0=[0.5, 0.7], p0=[0.0, 1.0, 1.0], tspan=(0.0, 12.0),
                         controls=[(1, 0.25 * np.arange(48), rng.uniform(0, 1, 48))], tunable_ic=[1, 2],
                         quadrature_indices=[], shooting_points=[0.0, 4.0, 8.0]), 1      # default tolerances
    else:
        spec, row = dict(model="lin1d", u0=[1.0], p0=[-2.0], tspan=(0.0, 1.0), controls=[], tunable_ic=[1],
                         quadrature_indices=[], shooting_points=[0.0, 0.5], kwargs=CORE), 1
    assert_ok(full_compare(spec, alg, 33, rng, objective_row=row))


@pytest.mark.parametrize("shooting", [None, (0.0, 3.0, 6.0, 9.0)])
def test_adaptive_oed_fisher_information_parity(shooting):
    """augmented Lotka OED at default tolerances, batch of designs: F, criteria and d x_end / d ps against the oracle"""
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    from oracle import pyoracle as po
    rng = np.random.default_rng(73)
    spec = P.lotka_oed_spec(shooting_points=None if shooting is None else list(shooting), fishing=rng.uniform(0, 1, 48))
    lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(adaptive=True))
    O = po.OracleLayer(spec)
    po.set_tolerances(1e-6, 1e-3)
    B = 6
    x0 = cb.to_vec(ps_t)
    if shooting is not None:
        x0 = O.forward_init(x0, method="tsit5_adaptive")
    X = np.tile(x0, (B, 1))
    X[:, -(24 + 79):] = rng.uniform(0, 1, (B, 24 + 79))            # the measurement weights of the last block(s)
    r = N.eval_host(X, nat.WANT_FIM | nat.WANT_XEND | nat.WANT_SENS)
    # at reltol = 1e-3 the controller's Float32 powers quantise the step size (a last-bit difference in the error
    # estimate moves a step by 6e-8 relative): two implementations agree to ~1e-5 reltol, not to 1e-10
    RTOL = 1e-8
    for b in range(B):
        o = O.eval(X[b], method="tsit5_adaptive")
        assert rel_err(r["xend"][b].reshape(O.I, O.nx).T, o["xend"]) <= RTOL
        F = o["last"][6:9]
        assert rel_err(r["fim"][b], [F[0], F[1], F[1], F[2]]) <= RTOL
    Jx, _, _ = O.jacobians(X[0], method="tsit5_adaptive")
    blocks = O.block_structure()
    off = 0
    for i in range(O.I):
        ns = blocks[i + 1] - blocks[i]
        S = r["sens"][0][off:off + O.nx * ns].reshape(ns, O.nx).T
        assert rel_err(S, Jx[O.nx * i:O.nx * (i + 1), blocks[i]:blocks[i + 1]]) <= RTOL
        off += O.nx * ns


def test_adaptive_mode_of_the_dense_model():
    """the 20-state model in the reference's default mode: the tensor-path kernel is fixed-step, the adaptive mode runs
    on the generic kernel (one thread per candidate, interval and direction) -- same parity bar"""
    import corleone_b200 as cb
    rng = np.random.default_rng(41)
    spec = dict(P.dense20_spec(3, rng), kwargs=dict(abstol=1e-9, reltol=1e-8))
    assert_ok(full_compare(spec, cb.Tsit5(adaptive=True), 3, rng, objective_row=3))


@pytest.mark.parametrize("name", ["vdp", "fuller", "hangglider", "ductedfan", "discrete_oed"])
def test_adaptive_parity_other_models(name):
    """forward-mode-AD models and an OED variant in adaptive mode"""
    import corleone_b200 as cb
    rng = np.random.default_rng(79)
    if name == "vdp":
        spec, row, scale = P.vdp_spec(5, rng), 3, 0.05
    elif name == "discrete_oed":
        spec = dict(P.lotka_oed_spec(shooting_points=[0.0, 3.0, 6.0, 9.0], fishing=rng.uniform(0, 1, 48)))
        # the discrete layer as the reference builds it (lib/CorleoneOED/src/oed.jl:87-113): [x; G] with the base parameters,
        # the measurement controls (p index 4, 5: beyond length(p)) stay in ps.controls without cutting the pieces, and
        # their grid points are saveat times -- sampled from the dense output of the adaptive steps
        spec.update(model="lotka2_oed_ds", u0=[0.5, 0.7] + [0.0] * 4, p0=[0.0, 1.0, 1.0],
                    saveat=np.union1d(spec["controls"][1][1], spec["controls"][2][1]))
        spec.pop("fim", None)
        row, scale = 1, 0.02
    else:
        spec, row, scale = P.bench_spec(name, 3 if name == "ductedfan" else 4)
    spec = dict(spec, kwargs=dict(abstol=1e-8, reltol=1e-6))
    from oracle import pyoracle as po
    O = po.OracleLayer(spec)
    ps = O.default_ps()[None, :] * (1.0 + scale * rng.standard_normal((5, O.nvars))) + scale * rng.standard_normal((5, O.nvars))
    assert_ok(full_compare(spec, cb.Tsit5(adaptive=True), 5, rng, objective_row=row, ps=ps))


def test_adaptive_failures_are_flagged_and_terminate():
    """a candidate whose integration cannot proceed (non-finite input; an explosive right-hand side that drives the
    step below resolution) ends with the status flag set instead of spinning; its neighbours are unaffected"""
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    lay, ps_t, st = _layer(dict(P.lotka_spec(), kwargs=CORE))
    N = lay._native(st, ps_t)
    X = np.tile(cb.to_vec(ps_t), (4, 1))
    X[1, 5] = np.nan
    X[2, N.block_structure()[1]] = 1e200          # node value: x' ~ -x y overflows at once
    r = N.eval_host(X, nat.WANT_XEND | nat.WANT_DEFECTS)
    assert list(r["status"] != 0) == [False, True, True, False]
    assert np.array_equal(r["xend"][0], r["xend"][3]) and np.isfinite(r["xend"][0]).all()
